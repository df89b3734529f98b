"""Command line of the engine, flag-compatible with the reference's entry point (main.py:13-79):

    python main.py -p {train,evaluate} --config_json FILE --gpu IDS [-v]

The reference opens a TensorFlow session and hands it to `image2label`; here the same class drives the B200
engine through the C ABI and the session argument is None.
"""
import argparse
import json
import os
import sys

PHASES = ("train", "evaluate")

# (flags, keyword arguments) of every option the reference accepts; names, destinations and defaults are its own
OPTIONS = (
    (("-v", "--verbose"), dict(dest="verbose", action="store_true", help="print the parsed arguments")),
    (("-p", "--phase"), dict(dest="phase", choices=list(PHASES), default=PHASES[0], metavar="[train evaluate]",
                             help="what to run (default: train)")),
    (("--config_json",), dict(dest="config_json", type=str, default="config.json", metavar="FILENAME",
                              help="model / training / evaluation configuration")),
    (("--gpu",), dict(dest="gpu", type=str, default="0", metavar="GPU_IDs",
                      help="value for CUDA_VISIBLE_DEVICES (default: 0); the engine runs on the first visible device")),
)


def parse_arguments(argv=None):
    parser = argparse.ArgumentParser(description="B200-native V-Net segmentation behind the vnet-tensorflow entry points")
    for flags, kwargs in OPTIONS:
        parser.add_argument(*flags, **kwargs)
    args = parser.parse_args(argv)
    if args.verbose:
        for name, value in sorted(vars(args).items()):
            print("{} = {}".format(name, value))
    return args


def run(args):
    os.environ["CUDA_VISIBLE_DEVICES"] = str(args.gpu)  # before the CUDA library is loaded, as main.py:62
    with open(args.config_json) as f:
        settings = json.load(f)
    from vnet_tensorflow_b200.model import image2label
    model = image2label(None, settings)
    if args.phase not in PHASES:
        sys.exit("Invalid training phase")
    getattr(model, args.phase)()


if __name__ == "__main__":
    run(parse_arguments())
