// vnb_infer -- native sliding-window inference driver over the C ABI of libvnet_b200.so.
//
// B200-native counterpart of the reference's only native component, cxx/tf_inference.cpp + main.cxx (TF-1.8 C++
// API + ITK, "Deprecated" in README.md:46,242): load the trained variables (cxx/tf_inference.cpp:96-144 runs the
// `nWeights` assigns of a frozen graph), window the intensities to 0..255 (:155-170), pad, run the overlapping patch
// grid (:218-274, :410-415) and average / arg-max the votes (:417-475), write the label volume.  Here the patch
// loop, the softmax accumulation and the arg-max run on the GPU behind vnb_evaluate_volume (model.py:866-937
// semantics: un-normalised softmax sums, first maximum wins) and this program only does file I/O -- no Python,
// no TensorFlow, no ITK.  The reference's B-spline resampling to 0.2 mm (:172-209) is data preparation and stays
// outside (out of scope).
//
//   vnb_infer --lib libvnet_b200.so --weights model.vnbw --image ct.nii [--image t2.nii ...] --out label.nii
//             --patch 64 64 64 --stride 32 32 32 --batch 2 --classes 2 [--labels 0 1] [--precision fp32|bf16x3|bf16]
//             [--window lo hi] [--channels 16 --levels 4 --convs 1 2 3 3 --bottom 3] [--device 0]
//
// Weight file (.vnbw, written by vnet_tensorflow_b200.checkpoint.export_binary): "VNBW" u32 version u32 count, then
// per variable: u32 name length, name bytes (TF variable name), u32 ndim, i64 dims[ndim], float32 data.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/vnet_b200.h"

namespace {

struct Api {
  void* so = nullptr;
  decltype(&vnb_create) create = nullptr;
  decltype(&vnb_destroy) destroy = nullptr;
  decltype(&vnb_set_param) set_param = nullptr;
  decltype(&vnb_evaluate_volume) evaluate_volume = nullptr;
  decltype(&vnb_last_error) last_error = nullptr;
  template <class F>
  void bind(F& f, const char* name) {
    f = reinterpret_cast<F>(dlsym(so, name));
    if (!f) throw std::runtime_error(std::string("missing symbol ") + name);
  }
  explicit Api(const std::string& path) {
    so = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!so) throw std::runtime_error(std::string("cannot load ") + path + ": " + dlerror());
    bind(create, "vnb_create");
    bind(destroy, "vnb_destroy");
    bind(set_param, "vnb_set_param");
    bind(evaluate_volume, "vnb_evaluate_volume");
    bind(last_error, "vnb_last_error");
  }
  void check(int rc, const char* what) const {
    if (rc != VNB_OK) throw std::runtime_error(std::string(what) + ": " + last_error());
  }
};

// ---- NIfTI-1, single file, uncompressed, little endian ------------------------------------------------------
struct Volume {
  int dim[3] = {0, 0, 0};
  float pixdim[3] = {1, 1, 1}, origin[3] = {0, 0, 0};
  float qfac = 1.f;          // pixdim[0]: handedness of the qform
  char orient[96] = {0};     // header bytes 252..347 verbatim: qform / sform codes, quaternion, offsets, srow_x/y/z,
                             // intent name -- written back unchanged so the label volume overlays the input
  bool has_orient = false;
  std::vector<float> data;   // file order: x fastest
};

template <class T>
void convert(const std::vector<char>& raw, size_t off, size_t n, float slope, float inter, std::vector<float>& out) {
  out.resize(n);
  const T* p = reinterpret_cast<const T*>(raw.data() + off);
  for (size_t i = 0; i < n; ++i) out[i] = static_cast<float>(p[i]) * slope + inter;
}

Volume read_nifti(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (raw.size() < 352) throw std::runtime_error(path + ": not a NIfTI-1 file");
  int32_t sizeof_hdr;
  std::memcpy(&sizeof_hdr, raw.data(), 4);
  if (sizeof_hdr != 348) throw std::runtime_error(path + ": only little-endian uncompressed NIfTI-1 (.nii) is supported");
  int16_t dim[8], datatype;
  float pixdim[8], vox_offset, slope, inter, qoff[3];
  std::memcpy(dim, raw.data() + 40, 16);
  std::memcpy(&datatype, raw.data() + 70, 2);
  std::memcpy(pixdim, raw.data() + 76, 32);
  std::memcpy(&vox_offset, raw.data() + 108, 4);
  std::memcpy(&slope, raw.data() + 112, 4);
  std::memcpy(&inter, raw.data() + 116, 4);
  std::memcpy(qoff, raw.data() + 268, 12);
  if (slope == 0.f) slope = 1.f;
  Volume v;
  for (int a = 0; a < 3; ++a) {
    v.dim[a] = dim[1 + a] > 0 ? dim[1 + a] : 1;
    v.pixdim[a] = pixdim[1 + a];
    v.origin[a] = qoff[a];
  }
  v.qfac = pixdim[0];
  std::memcpy(v.orient, raw.data() + 252, 92);
  v.has_orient = true;
  const size_t n = static_cast<size_t>(v.dim[0]) * v.dim[1] * v.dim[2];
  const size_t off = std::max<size_t>(static_cast<size_t>(vox_offset), 352);
  const size_t bytes[] = {1, 2, 4, 4, 8};
  int kind;
  switch (datatype) {
    case 2: kind = 0; break;     // uint8
    case 4: kind = 1; break;     // int16
    case 8: kind = 2; break;     // int32
    case 16: kind = 3; break;    // float32
    case 64: kind = 4; break;    // float64
    default: throw std::runtime_error(path + ": unsupported NIfTI datatype " + std::to_string(datatype));
  }
  if (raw.size() < off + n * bytes[kind]) throw std::runtime_error(path + ": truncated voxel data");
  switch (kind) {
    case 0: convert<uint8_t>(raw, off, n, slope, inter, v.data); break;
    case 1: convert<int16_t>(raw, off, n, slope, inter, v.data); break;
    case 2: convert<int32_t>(raw, off, n, slope, inter, v.data); break;
    case 3: convert<float>(raw, off, n, slope, inter, v.data); break;
    default: convert<double>(raw, off, n, slope, inter, v.data); break;
  }
  return v;
}

void write_nifti_i32(const std::string& path, const Volume& geom, const std::vector<int32_t>& data) {
  std::vector<char> hdr(352, 0);
  const int32_t sz = 348;
  std::memcpy(hdr.data(), &sz, 4);
  const int16_t dim[8] = {3, static_cast<int16_t>(geom.dim[0]), static_cast<int16_t>(geom.dim[1]), static_cast<int16_t>(geom.dim[2]), 1, 1, 1, 1};
  std::memcpy(hdr.data() + 40, dim, 16);
  const int16_t datatype = 8, bitpix = 32, qform = 1;
  std::memcpy(hdr.data() + 70, &datatype, 2);
  std::memcpy(hdr.data() + 72, &bitpix, 2);
  const float pixdim[8] = {geom.has_orient ? geom.qfac : 1.f, geom.pixdim[0], geom.pixdim[1], geom.pixdim[2], 1.f, 1.f, 1.f, 1.f};
  std::memcpy(hdr.data() + 76, pixdim, 32);
  const float vox_offset = 352.f, slope = 1.f, inter = 0.f;
  std::memcpy(hdr.data() + 108, &vox_offset, 4);
  std::memcpy(hdr.data() + 112, &slope, 4);
  std::memcpy(hdr.data() + 116, &inter, 4);
  if (geom.has_orient) {   // the input scan's qform / sform (direction, origin) carried through verbatim
    std::memcpy(hdr.data() + 252, geom.orient, 92);
  } else {
    std::memcpy(hdr.data() + 252, &qform, 2);
    std::memcpy(hdr.data() + 268, geom.origin, 12);
  }
  std::memcpy(hdr.data() + 344, "n+1", 4);
  std::ofstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot write " + path);
  f.write(hdr.data(), hdr.size());
  f.write(reinterpret_cast<const char*>(data.data()), data.size() * sizeof(int32_t));
}

// ---- weights ---------------------------------------------------------------------------------------------------
void load_weights(const Api& api, vnb_handle* h, const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  char magic[4];
  uint32_t version = 0, count = 0;
  f.read(magic, 4);
  f.read(reinterpret_cast<char*>(&version), 4);
  f.read(reinterpret_cast<char*>(&count), 4);
  if (!f || std::memcmp(magic, "VNBW", 4) != 0 || version != 1) throw std::runtime_error(path + ": not a VNBW version 1 file");
  std::vector<float> buf;
  for (uint32_t i = 0; i < count; ++i) {
    uint32_t len = 0, ndim = 0;
    f.read(reinterpret_cast<char*>(&len), 4);
    std::string name(len, '\0');
    f.read(&name[0], len);
    f.read(reinterpret_cast<char*>(&ndim), 4);
    size_t n = 1;
    for (uint32_t d = 0; d < ndim; ++d) {
      int64_t e = 0;
      f.read(reinterpret_cast<char*>(&e), 8);
      n *= static_cast<size_t>(e);
    }
    buf.resize(n);
    f.read(reinterpret_cast<char*>(buf.data()), n * sizeof(float));
    if (!f) throw std::runtime_error(path + ": truncated at variable " + name);
    api.check(api.set_param(h, name.c_str(), buf.data(), n * sizeof(float)), name.c_str());
  }
}

struct Args {
  std::string lib = "vnet_tensorflow_b200/libvnet_b200.so", weights, out;
  std::vector<std::string> images;
  int patch[3] = {64, 64, 64}, stride[3] = {32, 32, 32};
  int batch = 1, classes = 2, channels = 16, levels = 4, bottom = 3, device = 0, precision = VNB_PREC_BF16X3;
  std::vector<int> convs = {1, 2, 3, 3}, labels;
  bool window = false;
  float wlo = 0.f, whi = 255.f;
};

Args parse(int argc, char** argv) {
  Args a;
  auto need = [&](int i, int n) {
    if (i + n >= argc) throw std::runtime_error(std::string("missing value after ") + argv[i]);
  };
  for (int i = 1; i < argc; ++i) {
    const std::string k = argv[i];
    if (k == "--lib") { need(i, 1); a.lib = argv[++i]; }
    else if (k == "--weights") { need(i, 1); a.weights = argv[++i]; }
    else if (k == "--image") { need(i, 1); a.images.push_back(argv[++i]); }
    else if (k == "--out") { need(i, 1); a.out = argv[++i]; }
    else if (k == "--patch") { need(i, 3); for (int d = 0; d < 3; ++d) a.patch[d] = std::atoi(argv[++i]); }
    else if (k == "--stride") { need(i, 3); for (int d = 0; d < 3; ++d) a.stride[d] = std::atoi(argv[++i]); }
    else if (k == "--batch") { need(i, 1); a.batch = std::atoi(argv[++i]); }
    else if (k == "--classes") { need(i, 1); a.classes = std::atoi(argv[++i]); }
    else if (k == "--channels") { need(i, 1); a.channels = std::atoi(argv[++i]); }
    else if (k == "--levels") { need(i, 1); a.levels = std::atoi(argv[++i]); }
    else if (k == "--bottom") { need(i, 1); a.bottom = std::atoi(argv[++i]); }
    else if (k == "--device") { need(i, 1); a.device = std::atoi(argv[++i]); }
    else if (k == "--window") { need(i, 2); a.window = true; a.wlo = std::atof(argv[++i]); a.whi = std::atof(argv[++i]); }
    else if (k == "--precision") {
      need(i, 1);
      const std::string p = argv[++i];
      a.precision = p == "fp32" ? VNB_PREC_FP32 : p == "bf16" ? VNB_PREC_BF16 : VNB_PREC_BF16X3;
    } else if (k == "--convs") {
      a.convs.clear();
      while (i + 1 < argc && argv[i + 1][0] != '-') a.convs.push_back(std::atoi(argv[++i]));
    } else if (k == "--labels") {
      while (i + 1 < argc && (argv[i + 1][0] != '-' || std::isdigit(static_cast<unsigned char>(argv[i + 1][1])))) a.labels.push_back(std::atoi(argv[++i]));
    } else {
      throw std::runtime_error("unknown argument " + k);
    }
  }
  if (a.weights.empty() || a.images.empty() || a.out.empty()) throw std::runtime_error("--weights, --image and --out are required");
  if (static_cast<int>(a.convs.size()) != a.levels) throw std::runtime_error("--convs needs one entry per level");
  return a;
}

}  // namespace

int main(int argc, char** argv) {
  try {
    const Args a = parse(argc, argv);
    const Api api(a.lib);
    std::vector<Volume> ch;
    for (const std::string& p : a.images) ch.push_back(read_nifti(p));
    for (const Volume& v : ch)
      for (int d = 0; d < 3; ++d)
        if (v.dim[d] != ch[0].dim[d]) throw std::runtime_error("image channels differ in size");
    const int M = static_cast<int>(ch.size());
    const int* dim = ch[0].dim;
    // pad at the end of each axis up to the patch size (the reference pads with its Padding transform) and lay the
    // case out as [X][Y][Z][M] (NiftiDataset3D.py:150-158); intensities optionally windowed to 0..255
    int pd[3];
    for (int d = 0; d < 3; ++d) pd[d] = std::max(dim[d], a.patch[d]);
    std::vector<float> vol(static_cast<size_t>(pd[0]) * pd[1] * pd[2] * M, 0.f);
    for (int m = 0; m < M; ++m)
      for (int z = 0; z < dim[2]; ++z)
        for (int y = 0; y < dim[1]; ++y)
          for (int x = 0; x < dim[0]; ++x) {
            float v = ch[m].data[(static_cast<size_t>(z) * dim[1] + y) * dim[0] + x];
            if (a.window) v = (std::min(std::max(v, a.wlo), a.whi) - a.wlo) * (255.f / (a.whi - a.wlo));
            vol[((static_cast<size_t>(x) * pd[1] + y) * pd[2] + z) * M + m] = v;
          }
    vnb_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.in_channels = M;
    cfg.num_classes = a.classes;
    cfg.num_channels = a.channels;
    cfg.num_levels = a.levels;
    for (int l = 0; l < a.levels; ++l) cfg.num_convolutions[l] = a.convs[l];
    cfg.bottom_convolutions = a.bottom;
    for (int d = 0; d < 3; ++d) cfg.patch_shape[d] = a.patch[d];
    cfg.max_batch = a.batch;
    cfg.precision = a.precision;
    cfg.loss = VNB_LOSS_WEIGHTED_SORENSEN;
    for (int c = 0; c < a.classes && c < 8; ++c) cfg.loss_weights[c] = 1.f;
    cfg.loss_alpha = 1.f;
    cfg.optimizer = VNB_OPT_ADAM;
    cfg.learning_rate = 1e-2f;
    cfg.decay_factor = 0.99f;
    cfg.decay_steps = 100.f;
    cfg.momentum = 0.9f;
    vnb_handle* h = nullptr;
    api.check(api.create(&cfg, a.device, &h), "vnb_create");
    load_weights(api, h, a.weights);
    std::vector<int64_t> label(static_cast<size_t>(pd[0]) * pd[1] * pd[2]);
    const int32_t dims32[3] = {pd[0], pd[1], pd[2]}, stride32[3] = {a.stride[0], a.stride[1], a.stride[2]};
    api.check(api.evaluate_volume(h, vol.data(), dims32, stride32, a.batch, label.data(), nullptr, nullptr), "vnb_evaluate_volume");
    api.destroy(h);
    // crop the padding, back to file order; the reference writes the class index (model.py:934,945), --labels
    // (optional) writes that index's label value instead
    std::vector<int32_t> out(static_cast<size_t>(dim[0]) * dim[1] * dim[2]);
    for (int z = 0; z < dim[2]; ++z)
      for (int y = 0; y < dim[1]; ++y)
        for (int x = 0; x < dim[0]; ++x) {
          const int64_t c = label[(static_cast<size_t>(x) * pd[1] + y) * pd[2] + z];
          out[(static_cast<size_t>(z) * dim[1] + y) * dim[0] + x] =
              c < static_cast<int64_t>(a.labels.size()) ? a.labels[static_cast<size_t>(c)] : static_cast<int32_t>(c);
        }
    write_nifti_i32(a.out, ch[0], out);
    std::printf("vnb_infer: %dx%dx%d, %d channel(s) -> %s\n", dim[0], dim[1], dim[2], M, a.out.c_str());
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "vnb_infer: %s\n", e.what());
    return 1;
  }
}
