"""CLI entry point with the reference's flags (main.py:13-79):

    python main.py -p {train,evaluate} --config_json FILE --gpu IDS [-v]

The TF session of the reference is gone; `image2label(None, config)` drives the B200 engine instead.
"""
import argparse
import json
import os
import sys


def str2bool(v):
    return v.lower() in ("yes", "true", "t", "1")


def get_parser():
    parser = argparse.ArgumentParser(
        description='B200-native V-Net segmentation engine behind the vnet-tensorflow entry points.')
    parser.register('type', 'bool', str2bool)
    parser.add_argument('-v', '--verbose', dest='verbose', help='Show verbose output', action='store_true')
    parser.add_argument('-p', '--phase', dest='phase', help='Training phase (default= train)',
                        choices=['train', 'evaluate'], default='train', metavar='[train evaluate]')
    parser.add_argument('--config_json', dest='config_json', help='JSON file for model configuration', type=str,
                        default='config.json', metavar='FILENAME')
    parser.add_argument('--gpu', dest='gpu', default='0', type=str, help='Select GPU device(s) (default = 0)',
                        metavar='GPU_IDs')
    args = parser.parse_args()
    if args.verbose:
        for key in sorted(vars(args)):
            print("{} = {}".format(str(key), str(vars(args)[key])))
    return args


def main(args):
    os.environ["CUDA_VISIBLE_DEVICES"] = str(args.gpu)  # main.py:62
    with open(args.config_json) as config_json:
        config = json.load(config_json)
    from vnet_tensorflow_b200.model import image2label
    model = image2label(None, config)
    if args.phase == "train":
        model.train()
    elif args.phase == "evaluate":
        model.evaluate()
    else:
        sys.exit("Invalid training phase")


if __name__ == "__main__":
    main(get_parser())
