// C ABI of libvnet_b200.so (see include/vnet_b200.h).  Thin, exception-free shell over vnb::Engine.
#include "../../include/vnet_b200.h"

#include <memory>
#include <new>

#include "engine.cuh"
#include "conv_tc_impl.cuh"
#ifndef VNB_EMULATE
#include "comm.cuh"
#endif

struct vnb_handle {
  std::unique_ptr<vnb::Engine> engine;
  int device = 0;
#ifndef VNB_EMULATE
  std::unique_ptr<vnb::Comm> comm;
#endif
};

namespace {
thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
template <class F>
int guarded(F&& f) {
  try {
    f();
    return VNB_OK;
  } catch (const std::invalid_argument& e) {
    return fail(VNB_ERR_INVALID_ARG, e.what());
  } catch (const std::bad_alloc& e) {
    return fail(VNB_ERR_OOM, e.what());
  } catch (const std::exception& e) {
    const std::string m = e.what();
    if (m.find("out of memory") != std::string::npos) return fail(VNB_ERR_OOM, m);
    if (m.find("NCCL") != std::string::npos) return fail(VNB_ERR_NCCL, m);
    if (m.find("CUDA") != std::string::npos) return fail(VNB_ERR_CUDA, m);
    return fail(VNB_ERR_INTERNAL, m);
  } catch (...) {
    return fail(VNB_ERR_INTERNAL, "unknown exception");
  }
}
void need(const void* p, const char* what) {
  if (!p) throw std::invalid_argument(std::string(what) + " must not be NULL");
}
void select_device(vnb_handle* h) {
#ifndef VNB_EMULATE
  VNB_CUDA_OK(cudaSetDevice(h->device));
#else
  (void)h;
#endif
}
}  // namespace

extern "C" {

const char* vnb_last_error(void) { return g_last_error.c_str(); }
const char* vnb_version(void) {
#ifdef VNB_EMULATE
  return "vnet_b200 0.1 (CPU emulation build: tests only)";
#else
  return "vnet_b200 0.1 (sm_100a)";
#endif
}

int vnb_create(const vnb_config* c, int device, vnb_handle** out) {
  return guarded([&] {
    need(c, "cfg");
    need(out, "out");
    *out = nullptr;
#ifndef VNB_EMULATE
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
      throw std::runtime_error("CUDA: no CUDA device visible -- libvnet_b200 has no CPU fallback");
    if (device < 0 || device >= count) throw std::invalid_argument("device index out of range");
    cudaDeviceProp prop;
    VNB_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
      throw std::runtime_error("CUDA: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                               ", this library is built for sm_100a (B200) only");
    VNB_CUDA_OK(cudaSetDevice(device));
#endif
    vnb::EngineConfig e;
    e.in_channels = c->in_channels;
    e.num_classes = c->num_classes;
    e.num_channels = c->num_channels;
    e.num_levels = c->num_levels;
    for (int i = 0; i < 8; ++i) e.num_convolutions[i] = c->num_convolutions[i];
    e.bottom_convolutions = c->bottom_convolutions;
    for (int i = 0; i < 3; ++i) e.patch[i] = c->patch_shape[i];
    e.max_batch = c->max_batch;
    e.precision = c->precision;
    e.loss = c->loss;
    for (int i = 0; i < 8; ++i) e.loss_weights[i] = c->loss_weights[i];
    e.loss_alpha = c->loss_alpha;
    e.optimizer = c->optimizer;
    e.lr0 = c->learning_rate;
    e.decay_factor = c->decay_factor;
    e.decay_steps = c->decay_steps;
    e.momentum = c->momentum;
    e.flavour = c->graph_flavour;
    if (e.flavour < 0 || e.flavour > 1) throw std::invalid_argument("graph_flavour must be 0 (networks.VNet) or 1 (VNet.py)");
    if (e.precision < 0 || e.precision > 2) throw std::invalid_argument("precision must be VNB_PREC_*");
    if (e.loss < 0 || e.loss > VNB_LOSS_SORENSEN_FG) throw std::invalid_argument("loss must be VNB_LOSS_*");
    e.attention = c->attention;
    e.attention_loss = c->attention_loss;
    e.module_channels = c->module_channels > 0 ? c->module_channels : 64;
    if (e.attention < 0 || e.attention > 1) throw std::invalid_argument("attention must be 0 or 1");
    if (e.attention_loss < 0 || e.attention_loss > VNB_ATT_ABS) throw std::invalid_argument("attention_loss must be VNB_ATT_*");
    if (e.attention_loss != VNB_ATT_NONE && !e.attention) throw std::invalid_argument("attention_loss needs attention = 1");
    if (e.attention_loss == VNB_ATT_ABS && e.num_classes != 2)
      throw std::invalid_argument("attention_loss abs is defined for 2 classes only (train.py:394-398)");
    if (e.attention && e.num_classes < 2) throw std::invalid_argument("the attention path needs >= 2 classes");
    if (e.loss == VNB_LOSS_SORENSEN_FG && e.num_classes < 2) throw std::invalid_argument("sorensen_fg needs >= 2 classes");
    if (e.module_channels % 4 || e.module_channels > 256) throw std::invalid_argument("module_channels must be a multiple of 4, <= 256");
    if (e.optimizer < 0 || e.optimizer > VNB_OPT_NESTEROV) throw std::invalid_argument("optimizer must be VNB_OPT_*");
    for (int l = 0; l < e.num_levels && l < 8; ++l)
      if (e.num_convolutions[l] < 1) throw std::invalid_argument("NumConvolutions entries must be >= 1");
    if (e.bottom_convolutions < 1) throw std::invalid_argument("BottomConvolutions must be >= 1");
    std::unique_ptr<vnb_handle> h(new vnb_handle);
    h->device = device;
    h->engine.reset(new vnb::Engine(e));
    *out = h.release();
  });
}

int vnb_destroy(vnb_handle* h) {
  return guarded([&] {
    if (!h) return;
    select_device(h);
#ifndef VNB_EMULATE
    h->comm.reset();
#endif
    delete h;
  });
}

int vnb_num_params(vnb_handle* h, int* count) {
  return guarded([&] {
    need(h, "handle");
    need(count, "count");
    *count = static_cast<int>(h->engine->params().size());
  });
}

int vnb_param_info(vnb_handle* h, int index, const char** tf_name, int* ndim, int64_t dims[5], int* trainable) {
  return guarded([&] {
    need(h, "handle");
    const auto& ps = h->engine->params();
    if (index < 0 || index >= static_cast<int>(ps.size())) throw std::invalid_argument("parameter index out of range");
    const vnb::ParamEntry& e = ps[index];
    if (tf_name) *tf_name = e.name.c_str();
    if (ndim) *ndim = e.ndim;
    if (dims)
      for (int i = 0; i < 5; ++i) dims[i] = e.dims[i];
    if (trainable) *trainable = e.trainable ? 1 : 0;
  });
}

int vnb_set_slot(vnb_handle* h, const char* name, int slot, const void* host, size_t bytes) {
  return guarded([&] {
    need(h, "handle");
    need(name, "tf_name");
    need(host, "host");
    if (slot < 0 || slot > 3) throw std::invalid_argument("slot must be VNB_SLOT_*");
    select_device(h);
    h->engine->set_param(name, static_cast<const float*>(host), bytes, slot);
  });
}
int vnb_get_slot(vnb_handle* h, const char* name, int slot, void* host, size_t bytes) {
  return guarded([&] {
    need(h, "handle");
    need(name, "tf_name");
    need(host, "host");
    if (slot < 0 || slot > 3) throw std::invalid_argument("slot must be VNB_SLOT_*");
    select_device(h);
    h->engine->get_param(name, static_cast<float*>(host), bytes, slot);
  });
}
int vnb_set_param(vnb_handle* h, const char* name, const void* host, size_t bytes) {
  return vnb_set_slot(h, name, VNB_SLOT_VALUE, host, bytes);
}
int vnb_get_param(vnb_handle* h, const char* name, void* host, size_t bytes) {
  return vnb_get_slot(h, name, VNB_SLOT_VALUE, host, bytes);
}
int vnb_get_step(vnb_handle* h, int64_t* s) {
  return guarded([&] {
    need(h, "handle");
    need(s, "global_step");
    *s = h->engine->global_step();
  });
}
int vnb_set_step(vnb_handle* h, int64_t s) {
  return guarded([&] {
    need(h, "handle");
    if (s < 0) throw std::invalid_argument("global_step must be >= 0");
    h->engine->set_global_step(s);
  });
}

int vnb_forward(vnb_handle* h, const float* images, int n, float* logits, float* softmax, int64_t* argmax) {
  return guarded([&] {
    need(h, "handle");
    need(images, "images");
    select_device(h);
    h->engine->forward_host(images, n, logits, softmax, reinterpret_cast<long long*>(argmax));
  });
}

int vnb_evaluate_volume(vnb_handle* h, const float* volume, const int32_t dims[3], const int32_t stride[3], int batch,
                        int64_t* label, float* softmax_sum, float* weight) {
  return guarded([&] {
    need(h, "handle");
    need(volume, "volume");
    need(dims, "dims");
    need(stride, "stride");
    select_device(h);
    const int d[3] = {dims[0], dims[1], dims[2]}, st[3] = {stride[0], stride[1], stride[2]};
    h->engine->evaluate_volume_host(volume, d, st, batch, reinterpret_cast<long long*>(label), softmax_sum, weight);
  });
}

int vnb_loss(vnb_handle* h, const float* images, const int32_t* labels, int n, float* loss_out, double* terms) {
  return guarded([&] {
    need(h, "handle");
    need(images, "images");
    need(labels, "labels");
    select_device(h);
    const float l = h->engine->loss_host(images, labels, n, terms);
    if (loss_out) *loss_out = l;
  });
}

int vnb_forward_backward(vnb_handle* h, const float* images, const int32_t* labels, int n, float dropout,
                         uint64_t seed, int update_moving, float* loss_out) {
  return guarded([&] {
    need(h, "handle");
    need(images, "images");
    need(labels, "labels");
    if (!(dropout >= 0.f && dropout < 1.f)) throw std::invalid_argument("dropout_rate must be in [0,1)");
    select_device(h);
    h->engine->upload_batch(images, labels, n);
#ifndef VNB_EMULATE
    if (h->comm) h->comm->begin_step(*h->engine);
#endif
    h->engine->forward_backward_device(n, dropout, seed, update_moving != 0);
    if (loss_out) *loss_out = h->engine->read_loss();
  });
}

int vnb_apply_gradients(vnb_handle* h) {
  return guarded([&] {
    need(h, "handle");
    select_device(h);
    int world = 1;
#ifndef VNB_EMULATE
    if (h->comm) {
      h->comm->finish_allreduce(*h->engine);
      world = h->comm->world();
    }
#endif
    h->engine->optimizer_step(world);
  });
}

int vnb_train_step(vnb_handle* h, const float* images, const int32_t* labels, int n, float dropout, uint64_t seed,
                   float* loss_out) {
  int rc = vnb_forward_backward(h, images, labels, n, dropout, seed, 1, nullptr);
  if (rc != VNB_OK) return rc;
  rc = vnb_apply_gradients(h);
  if (rc != VNB_OK) return rc;
  if (loss_out) return guarded([&] { *loss_out = h->engine->read_loss(); });
  return VNB_OK;
}

int vnb_comm_unique_id(void* id_out) {
  return guarded([&] {
    need(id_out, "id_out");
#ifndef VNB_EMULATE
    vnb::Comm::unique_id(id_out);
#else
    throw std::invalid_argument("emulation build has no communicator");
#endif
  });
}
int vnb_comm_init(vnb_handle* h, int rank, int world, const void* uid) {
  return guarded([&] {
    need(h, "handle");
    need(uid, "unique_id");
    if (world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("bad rank/world");
#ifndef VNB_EMULATE
    select_device(h);
    h->comm.reset(new vnb::Comm(rank, world, uid, *h->engine));
#else
    throw std::invalid_argument("emulation build has no communicator");
#endif
  });
}
int vnb_comm_world(vnb_handle* h, int* rank, int* world) {
  return guarded([&] {
    need(h, "handle");
    int r = 0, w = 1;
#ifndef VNB_EMULATE
    if (h->comm) {
      r = h->comm->rank();
      w = h->comm->world();
    }
#endif
    if (rank) *rank = r;
    if (world) *world = w;
  });
}

int vnb_comm_sync_bn(vnb_handle* h, int on) {
  return guarded([&] {
    need(h, "handle");
#ifndef VNB_EMULATE
    select_device(h);
    if (!h->comm) throw std::invalid_argument("vnb_comm_sync_bn needs vnb_comm_init first");
    h->comm->enable_sync_bn(*h->engine, on != 0);
#else
    (void)on;
    throw std::runtime_error("the NCCL communicator is not part of the emulation build; use vnb_set_stats_allreduce");
#endif
  });
}

int vnb_set_stats_allreduce(vnb_handle* h, vnb_allreduce_fn fn, void* user, int world) {
  return guarded([&] {
    need(h, "handle");
    if (!fn) {
      h->engine->set_stats_hook(nullptr, 1);
      return;
    }
    if (world < 1) throw std::invalid_argument("world must be >= 1");
    vnb::Engine* e = h->engine.get();
    e->set_stats_hook(
        [e, fn, user](double* dev, int n) {  // host round trip: the exchange itself belongs to the caller
          std::vector<double> host(static_cast<size_t>(n));
          VNB_CUDA_OK(cudaMemcpyAsync(host.data(), dev, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream()));
          VNB_CUDA_OK(cudaStreamSynchronize(e->stream()));
          if (fn(host.data(), n, user) != 0) throw std::runtime_error("statistics all-reduce callback failed");
          VNB_CUDA_OK(cudaMemcpyAsync(dev, host.data(), sizeof(double) * n, cudaMemcpyHostToDevice, e->stream()));
          VNB_CUDA_OK(cudaStreamSynchronize(e->stream()));
        },
        world);
  });
}

int vnb_upload_batch(vnb_handle* h, const float* images, const int32_t* labels, int n) {
  return guarded([&] {
    need(h, "handle");
    need(images, "images");
    need(labels, "labels");
    select_device(h);
    h->engine->upload_batch(images, labels, n);
    h->engine->sync();
  });
}
int vnb_train_step_resident(vnb_handle* h, int n, float dropout, uint64_t seed) {
  return guarded([&] {
    need(h, "handle");
    if (n < 1 || n > h->engine->config().max_batch) throw std::invalid_argument("batch size outside [1, max_batch]");
    select_device(h);
    int world = 1;
#ifndef VNB_EMULATE
    if (h->comm) h->comm->begin_step(*h->engine);
#endif
    h->engine->forward_backward_device(n, dropout, seed, true);
#ifndef VNB_EMULATE
    if (h->comm) {
      h->comm->finish_allreduce(*h->engine);
      world = h->comm->world();
    }
#endif
    h->engine->optimizer_step(world);
  });
}
int vnb_stage_batch(vnb_handle* h, const float* images, const int32_t* labels, int n) {
  return guarded([&] {
    need(h, "handle");
    need(images, "images");
    need(labels, "labels");
    select_device(h);
    h->engine->stage_batch(images, labels, n);
  });
}
int vnb_train_step_staged(vnb_handle* h, float dropout, uint64_t seed, float* loss_out) {
  int n = 0;
  int rc = guarded([&] {
    need(h, "handle");
    select_device(h);
    n = h->engine->commit_staged();
  });
  if (rc != VNB_OK) return rc;
  rc = vnb_train_step_resident(h, n, dropout, seed);
  if (rc != VNB_OK) return rc;
  if (loss_out) return guarded([&] { *loss_out = h->engine->read_loss(); });
  return VNB_OK;
}
int vnb_host_alloc(vnb_handle* h, size_t bytes, void** out) {
  return guarded([&] {
    need(h, "handle");
    need(out, "out");
    *out = nullptr;
    select_device(h);   // page-locked through the handle's own context (no stray context on device 0)
    VNB_CUDA_OK(cudaMallocHost(out, bytes ? bytes : 1));
  });
}
int vnb_host_free(void* p) {
  return guarded([&] {
    if (p) VNB_CUDA_OK(cudaFreeHost(p));
  });
}
int vnb_event_record(vnb_handle* h, int which) {
  return guarded([&] {
    need(h, "handle");
    select_device(h);
    h->engine->event_record(which);
  });
}
int vnb_event_elapsed_ms(vnb_handle* h, float* ms) {
  return guarded([&] {
    need(h, "handle");
    need(ms, "ms");
    select_device(h);
    *ms = h->engine->event_elapsed_ms();
  });
}
int vnb_profile_enable(vnb_handle* h, int on) {
  return guarded([&] {
    need(h, "handle");
    h->engine->profile_enable(on != 0);
  });
}
int vnb_profile_read(vnb_handle* h, int cls, double* ms, int64_t* launches, double* flops) {
  return guarded([&] {
    need(h, "handle");
    need(ms, "ms");
    need(launches, "launches");
    need(flops, "flops");
    select_device(h);
    long long n = 0;
    h->engine->profile_read(cls, ms, &n, flops);
    *launches = n;
  });
}
int vnb_read_metrics(vnb_handle* h, int n, uint64_t* confusion, uint64_t* auc_hist) {
  return guarded([&] {
    need(h, "handle");
    need(confusion, "confusion");
    select_device(h);
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t) && VNB_AUC_BINS == vnb::kAucBins, "metrics layout");
    h->engine->read_metrics(n, reinterpret_cast<unsigned long long*>(confusion), reinterpret_cast<unsigned long long*>(auc_hist));
  });
}
int vnb_profile_count(vnb_handle* h, int64_t* launches) {
  return guarded([&] {
    need(h, "handle");
    need(launches, "launches");
    *launches = static_cast<int64_t>(h->engine->profile_count());
  });
}
int vnb_profile_launch(vnb_handle* h, int64_t index, int* cls, double* ms, double* flops, char* label, size_t label_bytes) {
  return guarded([&] {
    need(h, "handle");
    need(cls, "kernel_class");
    need(ms, "ms");
    need(flops, "flops");
    select_device(h);
    std::string text;
    if (index < 0 || !h->engine->profile_launch(static_cast<size_t>(index), cls, ms, flops, &text))
      throw std::invalid_argument("profile record index out of range");
    if (label && label_bytes) {
      const size_t n = std::min(text.size(), label_bytes - 1);
      memcpy(label, text.data(), n);
      label[n] = 0;
    }
  });
}

int vnb_set_distmap(vnb_handle* h, const float* distmap, int n) {
  return guarded([&] {
    need(h, "handle");
    need(distmap, "distmap");
    select_device(h);
    h->engine->set_distmap(distmap, n);
  });
}
int vnb_read_losses(vnb_handle* h, float out[3]) {
  return guarded([&] {
    need(h, "handle");
    need(out, "out");
    select_device(h);
    h->engine->read_losses(out);
  });
}
int vnb_read_loss_parts(vnb_handle* h, float out[2]) {
  return guarded([&] {
    need(h, "handle");
    need(out, "out");
    select_device(h);
    h->engine->read_loss_parts(out);
  });
}
int vnb_read_softmax_attention(vnb_handle* h, float* host, size_t bytes, int n) {
  return guarded([&] {
    need(h, "handle");
    need(host, "host");
    select_device(h);
    const vnb::EngineConfig& c = h->engine->config();
    if (n < 1 || n > c.max_batch) throw std::invalid_argument("batch size outside [1, max_batch]");
    if (bytes != sizeof(float) * n * (size_t)c.patch[0] * c.patch[1] * c.patch[2] * c.num_classes)
      throw std::invalid_argument("read_softmax_attention size mismatch");
    h->engine->read_softmax_attention(host, n);
  });
}
int vnb_sync(vnb_handle* h) {
  return guarded([&] {
    need(h, "handle");
    select_device(h);
    h->engine->sync();
  });
}
int vnb_gpu_launches(vnb_handle* h, int64_t* count) {
  return guarded([&] {
    need(h, "handle");
    need(count, "count");
    *count = h->engine->gpu_launches();
  });
}
int vnb_read_tensor(vnb_handle* h, const char* scope, int kind, float* host, size_t bytes, int n) {
  return guarded([&] {
    need(h, "handle");
    need(scope, "scope");
    need(host, "host");
    select_device(h);
    h->engine->read_tensor(scope, kind, host, bytes, n);
  });
}

}  // extern "C"

// ---- standalone convolution ops (test hooks) ------------------------------------------------------
namespace {
struct DevBuf {
  void* p = nullptr;
  explicit DevBuf(size_t bytes) {
    if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) throw std::runtime_error("CUDA: out of memory in op hook");
  }
  ~DevBuf() { cudaFree(p); }
  template <class T>
  T* as() { return static_cast<T*>(p); }
};
void op_device(int device) {
#ifndef VNB_EMULATE
  VNB_CUDA_OK(cudaSetDevice(device));
#else
  (void)device;
#endif
}
template <int KS>
void launch_conv_ref(const vnb::Conv5Args& p, int co) {
  using namespace vnb;
#ifndef VNB_EMULATE
  VNB_CUDA_OK(cudaFuncSetAttribute(conv_ref_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ConvRefGeom<KS>::SMEM));
#endif
  const Dims& dims = p.dims;
  const int tiles = ((dims.W + kC5_TW - 1) / kC5_TW) * ((dims.H + kC5_TH - 1) / kC5_TH) * ((dims.D + kC5_TD - 1) / kC5_TD);
  dim3 grid(tiles, (co + kC5_CO - 1) / kC5_CO, p.N);
  VNB_LAUNCH(conv_ref_kernel<KS>, grid, 256, ConvRefGeom<KS>::SMEM, 0, p);
}
template <int KS>
void launch_wgrad_ref(const vnb::Wgrad5Args& a, dim3 grid) {
  using namespace vnb;
#ifndef VNB_EMULATE
  VNB_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_ref_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgradRefGeom<KS>::SMEM));
#endif
  VNB_LAUNCH(conv_wgrad_ref_kernel<KS>, grid, 256, WgradRefGeom<KS>::SMEM, 0, a);
}

void op_conv(int ks, int precision, const float* x, const float* w, const float* bias, const float* res, float* y, int n,
             vnb::Dims dims, int cin, int cout, bool dgrad_form) {
  using namespace vnb;
  const size_t V = static_cast<size_t>(n) * dims.D * dims.H * dims.W;
  const int ci = dgrad_form ? cout : cin, co = dgrad_form ? cin : cout;  // channels seen by the kernel
  const size_t wbytes = static_cast<size_t>(ks * ks * ks) * cin * cout * 4;
  DevBuf dx(V * ci * 4), dw(wbytes), dwf(wbytes), dy(V * co * 4), db(co * 4), dr(V * co * 4);
  VNB_CUDA_OK(cudaMemcpy(dx.p, x, V * ci * 4, cudaMemcpyHostToDevice));
  VNB_CUDA_OK(cudaMemcpy(dw.p, w, wbytes, cudaMemcpyHostToDevice));
  if (bias) VNB_CUDA_OK(cudaMemcpy(db.p, bias, co * 4, cudaMemcpyHostToDevice));
  if (res) VNB_CUDA_OK(cudaMemcpy(dr.p, res, V * co * 4, cudaMemcpyHostToDevice));
  if (precision != VNB_PREC_FP32) {
    tc_op_conv5(precision, dx.as<float>(), dw.as<float>(), bias ? db.as<float>() : nullptr, res ? dr.as<float>() : nullptr,
                dy.as<float>(), n, dims, cin, cout, dgrad_form, ks);
    VNB_CUDA_OK(cudaDeviceSynchronize());
    VNB_CUDA_OK(cudaMemcpy(y, dy.p, V * co * 4, cudaMemcpyDeviceToHost));
    return;
  }
  const float* wk = dw.as<float>();
  if (dgrad_form) {
    VNB_LAUNCH(flip_transpose_w_kernel, 1024, 256, 0, 0, (const float*)dw.as<float>(), dwf.as<float>(), cin, cout, ks * ks * ks);
    wk = dwf.as<float>();
  }
  Conv5Args p;
  p.in1 = dx.as<float>();
  p.in2 = nullptr;
  p.C1 = ci;
  p.C2 = 0;
  p.w = wk;
  p.bias = bias ? db.as<float>() : nullptr;
  p.res = res ? dr.as<float>() : nullptr;
  p.out1 = dy.as<float>();
  p.out2 = nullptr;
  p.Co1 = co;
  p.Co2 = 0;
  p.acc1 = p.acc2 = 0;
  p.dims = dims;
  p.N = n;
  if (ks == 3) launch_conv_ref<3>(p, co); else launch_conv_ref<5>(p, co);
  VNB_CUDA_OK(cudaDeviceSynchronize());
  VNB_CUDA_OK(cudaGetLastError());
  VNB_CUDA_OK(cudaMemcpy(y, dy.p, V * co * 4, cudaMemcpyDeviceToHost));
}

void op_wgrad(int ks, int precision, const float* x, const float* dy, float* dw, int n, int d, int h, int w_, int cin, int cout) {
  using namespace vnb;
  const Dims dims{d, h, w_};
  const size_t V = static_cast<size_t>(n) * d * h * w_;
  const size_t wbytes = static_cast<size_t>(ks * ks * ks) * cin * cout * 4;
  DevBuf bx(V * cin * 4), bdy(V * cout * 4), bdw(wbytes);
  VNB_CUDA_OK(cudaMemcpy(bx.p, x, V * cin * 4, cudaMemcpyHostToDevice));
  VNB_CUDA_OK(cudaMemcpy(bdy.p, dy, V * cout * 4, cudaMemcpyHostToDevice));
  VNB_CUDA_OK(cudaMemset(bdw.p, 0, wbytes));
  if (precision != VNB_PREC_FP32) {
    tc_op_wgrad5(precision, bx.as<float>(), bdy.as<float>(), bdw.as<float>(), n, dims, cin, cout, ks);
    VNB_CUDA_OK(cudaDeviceSynchronize());
    VNB_CUDA_OK(cudaMemcpy(dw, bdw.p, wbytes, cudaMemcpyDeviceToHost));
    return;
  }
  Wgrad5Args a;
  a.in1 = bx.as<float>();
  a.in2 = nullptr;
  a.C1 = cin;
  a.C2 = 0;
  a.dz = bdy.as<float>();
  a.Cout = cout;
  a.dw = bdw.as<float>();
  a.dims = dims;
  a.N = n;
  const long long ntiles = static_cast<long long>((w_ + kW5_TW - 1) / kW5_TW) * ((h + kW5_TH - 1) / kW5_TH) * ((d + kW5_TD - 1) / kW5_TD) * n;
  const int pairs = ((cin + 15) / 16) * ((cout + 15) / 16);
  a.tiles_per_block = static_cast<int>(std::max<long long>(1, ntiles / 64));
  const long long splits = (ntiles + a.tiles_per_block - 1) / a.tiles_per_block;
  dim3 grid(static_cast<unsigned>(splits), pairs);
  if (ks == 3) launch_wgrad_ref<3>(a, grid); else launch_wgrad_ref<5>(a, grid);
  VNB_CUDA_OK(cudaDeviceSynchronize());
  VNB_CUDA_OK(cudaGetLastError());
  VNB_CUDA_OK(cudaMemcpy(dw, bdw.p, wbytes, cudaMemcpyDeviceToHost));
}
}  // namespace

extern "C" {

#define VNB_DEFINE_CONV_OPS(KS)                                                                                      \
  int vnb_op_conv##KS##_fprop(int device, int precision, const float* x, const float* w, const float* bias,           \
                              const float* residual, float* y, int n, int d, int h, int w_, int cin, int cout) {      \
    return guarded([&] {                                                                                              \
      need(x, "x");                                                                                                   \
      need(w, "w");                                                                                                   \
      need(y, "y");                                                                                                   \
      op_device(device);                                                                                              \
      op_conv(KS, precision, x, w, bias, residual, y, n, vnb::Dims{d, h, w_}, cin, cout, false);                      \
    });                                                                                                               \
  }                                                                                                                   \
  int vnb_op_conv##KS##_dgrad(int device, int precision, const float* dy, const float* w, float* dx, int n, int d,    \
                              int h, int w_, int cin, int cout) {                                                     \
    return guarded([&] {                                                                                              \
      need(dy, "dy");                                                                                                 \
      need(w, "w");                                                                                                   \
      need(dx, "dx");                                                                                                 \
      op_device(device);                                                                                              \
      op_conv(KS, precision, dy, w, nullptr, nullptr, dx, n, vnb::Dims{d, h, w_}, cin, cout, true);                   \
    });                                                                                                               \
  }                                                                                                                   \
  int vnb_op_conv##KS##_wgrad(int device, int precision, const float* x, const float* dy, float* dw, int n, int d,    \
                              int h, int w_, int cin, int cout) {                                                     \
    return guarded([&] {                                                                                              \
      need(x, "x");                                                                                                   \
      need(dy, "dy");                                                                                                 \
      need(dw, "dw");                                                                                                 \
      op_device(device);                                                                                              \
      op_wgrad(KS, precision, x, dy, dw, n, d, h, w_, cin, cout);                                                     \
    });                                                                                                               \
  }
VNB_DEFINE_CONV_OPS(5)
VNB_DEFINE_CONV_OPS(3)
#undef VNB_DEFINE_CONV_OPS

}  // extern "C"

// ---- per-op hooks of the remaining hot-path kernels: host buffers in, host buffers out -----------------
namespace {
struct HostToDev {   // device copy of a host array (nullptr stays nullptr)
  DevBuf buf;
  HostToDev(const void* host, size_t bytes) : buf(bytes) {
    if (host) VNB_CUDA_OK(cudaMemcpy(buf.p, host, bytes, cudaMemcpyHostToDevice));
    else VNB_CUDA_OK(cudaMemset(buf.p, 0, bytes ? bytes : 16));
  }
};
void op_sync_and_check() {
  VNB_CUDA_OK(cudaDeviceSynchronize());
  VNB_CUDA_OK(cudaGetLastError());
}
inline int op_blocks(long long n, int block, int cap) {
  return static_cast<int>(std::max<long long>(1, std::min<long long>((n + block - 1) / block, cap)));
}
}  // namespace

extern "C" {

/* 2x2x2 stride-2 kernels (layers2.py:65-94), op: 0 down fprop / up dgrad (gather), 1 up fprop / down dgrad (scatter),
 * 2 filter gradient.  fine [n][2dc][2hc][2wc][cf], coarse [n][dc][hc][wc][cc], w and dw [2][2][2][cf][cc]. */
int vnb_op_k2(int device, int precision, int op, const float* fine, const float* coarse, const float* w, const float* bias,
              float* out, int n, int dc, int hc, int wc, int cf, int cc) {
  return guarded([&] {
    using namespace vnb;
    if (op != 2) need(w, "w");
    need(out, "out");
    if (op < 0 || op > 2) throw std::invalid_argument("vnb_op_k2: op must be 0 (gather), 1 (scatter) or 2 (filter gradient)");
    op_device(device);
    const long long M = static_cast<long long>(n) * dc * hc * wc;
    const size_t fine_b = static_cast<size_t>(M) * 8 * cf * 4, coarse_b = static_cast<size_t>(M) * cc * 4, w_b = static_cast<size_t>(8) * cf * cc * 4;
    HostToDev dfine(op != 1 ? fine : nullptr, fine_b), dcoarse(op != 0 ? coarse : nullptr, coarse_b), dw_(op != 2 ? w : nullptr, w_b);
    HostToDev db(bias, static_cast<size_t>(op == 0 ? cc : cf) * 4);
    K2Args p{};
    p.fine_in = dfine.buf.as<float>();
    p.coarse_in = dcoarse.buf.as<float>();
    p.fine_out = dfine.buf.as<float>();
    p.coarse_out = dcoarse.buf.as<float>();
    p.w = dw_.buf.as<float>();
    p.dw = dw_.buf.as<float>();
    p.bias = bias ? db.buf.as<float>() : nullptr;
    p.CF = cf;
    p.CC = cc;
    p.cd = Dims{dc, hc, wc};
    p.N = n;
    p.accumulate = 0;
    const bool tiled = cf % 16 == 0 && cc % 16 == 0, mma = precision != VNB_PREC_FP32;
    if (op == 0) {
      if (mma && tc_op_k2(false, dfine.buf.as<float>(), dcoarse.buf.as<float>(), p.w, p.bias, n, p.cd, cf, cc)) {
        // tcgen05 + TMA gather (k2_tc.cuh)
      } else if (tiled) {
        dim3 grid(static_cast<unsigned>((M + kK2_BM - 1) / kK2_BM), (cc + kK2_BN - 1) / kK2_BN);
        if (mma) VNB_LAUNCH(k2_gather_mma_kernel, grid, 256, 0, 0, p, M);
        else VNB_LAUNCH(k2_gather_tiled_kernel<false>, grid, 256, 0, 0, p, M);
      } else {
        VNB_LAUNCH(k2_gather_kernel, op_blocks(M * cc, 256, 1 << 20), 256, 0, 0, p);
      }
      op_sync_and_check();
      VNB_CUDA_OK(cudaMemcpy(out, dcoarse.buf.p, coarse_b, cudaMemcpyDeviceToHost));
    } else if (op == 1) {
      if (mma && tc_op_k2(true, dfine.buf.as<float>(), dcoarse.buf.as<float>(), p.w, p.bias, n, p.cd, cf, cc)) {
        // tcgen05 + TMA depth-to-space scatter (k2_tc.cuh)
      } else if (tiled) {
        dim3 grid(static_cast<unsigned>((M + kK2_BM - 1) / kK2_BM), (8 * cf + kK2_BN - 1) / kK2_BN);
        if (mma) VNB_LAUNCH(k2_scatter_mma_kernel, grid, 256, 0, 0, p, M);
        else VNB_LAUNCH(k2_scatter_tiled_kernel<false>, grid, 256, 0, 0, p, M);
      } else {
        VNB_LAUNCH(k2_scatter_kernel, op_blocks(M * 8 * cf, 256, 1 << 20), 256, 0, 0, p);
      }
      op_sync_and_check();
      VNB_CUDA_OK(cudaMemcpy(out, dfine.buf.p, fine_b, cudaMemcpyDeviceToHost));
    } else {
      VNB_CUDA_OK(cudaMemset(dw_.buf.p, 0, w_b));
      if (mma && tc_op_k2_wgrad(p.fine_in, p.coarse_in, p.dw, n, p.cd, cf, cc)) {
        // tcgen05 filter gradient (k2_tc.cuh)
      } else if (tiled) {
        const int gx = (8 * cf + kK2_BM - 1) / kK2_BM, gy = (cc + kK2_BN - 1) / kK2_BN;
        long long splits = std::max<long long>(1, std::min<long long>((M + 255) / 256, (4 * 148 + gx * gy - 1) / (gx * gy)));
        const long long mps = ((M + splits - 1) / splits + kK2_BK - 1) / kK2_BK * kK2_BK;
        splits = (M + mps - 1) / mps;
        dim3 grid(gx, gy, static_cast<unsigned>(splits));
        if (mma) VNB_LAUNCH(k2_wgrad_mma_kernel, grid, 256, 0, 0, p, M, mps);
        else VNB_LAUNCH(k2_wgrad_tiled_kernel<false>, grid, 256, 0, 0, p, M, mps);
      } else {
        const long long outs = 8LL * cf * cc;
        const int oblocks = static_cast<int>((outs + 255) / 256);
        long long splits = std::max<long long>(1, std::min<long long>(M, (2 * 1184 + oblocks - 1) / oblocks));
        const int vps = static_cast<int>((M + splits - 1) / splits);
        splits = (M + vps - 1) / vps;
        dim3 grid(oblocks, static_cast<unsigned>(splits));
        VNB_LAUNCH(k2_wgrad_kernel, grid, 256, 0, 0, p, vps);
      }
      op_sync_and_check();
      VNB_CUDA_OK(cudaMemcpy(out, dw_.buf.p, w_b, cudaMemcpyDeviceToHost));
    }
  });
}

/* Training-mode batch norm + PReLU (networks.py:319, layers2.py:97-99) on a [V][C] tensor, plain chain:
 * y = prelu(gamma * (z - mean) / sqrt(var + 1e-3) + beta); alpha may be null (no activation). */
int vnb_op_bn_fwd(int device, const float* z, const float* gamma, const float* beta, const float* alpha, float* y,
                  double* mean_out, double* var_out, long long voxels, int c) {
  return guarded([&] {
    using namespace vnb;
    need(z, "z"); need(gamma, "gamma"); need(beta, "beta"); need(y, "y");
    if (c % 4 || 256 % (c / 4)) throw std::invalid_argument("vnb_op_bn_fwd: C must be a multiple of 4 with C/4 dividing 256");
    op_device(device);
    const size_t nb = static_cast<size_t>(voxels) * c * 4;
    HostToDev dz(z, nb), dg(gamma, c * 4), dbt(beta, c * 4), da(alpha, c * 4), dmm(nullptr, c * 4), dmv(nullptr, c * 4);
    DevBuf dy(nb), partial(static_cast<size_t>(kMaxRedBlocks) * 2 * c * 8), mean(c * 8), var(c * 8), scale(c * 4), shift(c * 4);
    const unsigned total4 = static_cast<unsigned>(voxels * c / 4);
    const int nblk = op_blocks(total4, 256 * 8, kMaxRedBlocks / 2);
    VNB_LAUNCH(bn_stats_v4_kernel, nblk, 256, 0, 0, (const float*)dz.buf.as<float>(), c, total4, partial.as<double>());
    BnParams bp{};
    bp.gamma[0] = dg.buf.as<float>();
    bp.beta[0] = dbt.buf.as<float>();
    bp.moving_mean[0] = dmm.buf.as<float>();
    bp.moving_var[0] = dmv.buf.as<float>();
    VNB_LAUNCH(bn_finalize_fwd_kernel, c, 128, 0, 0, (const double*)partial.as<double>(), nblk, 2, c, static_cast<double>(voxels), (int)CH_S, bp,
               0, 0, mean.as<double>(), var.as<double>(), scale.as<float>(), shift.as<float>(), 0);
    ApplyArgs ap{};
    ap.z = dz.buf.as<float>();
    ap.a = dy.as<float>();
    ap.scale = scale.as<float>();
    ap.shift = shift.as<float>();
    ap.alpha = alpha ? da.buf.as<float>() : nullptr;
    ap.total = voxels * c;
    ap.C = c;
    VNB_LAUNCH(bn_apply_v4_kernel, op_blocks(ap.total / 4, 256, 2 * kMaxRedBlocks), 256, 0, 0, ap);
    op_sync_and_check();
    VNB_CUDA_OK(cudaMemcpy(y, dy.p, nb, cudaMemcpyDeviceToHost));
    if (mean_out) VNB_CUDA_OK(cudaMemcpy(mean_out, mean.p, c * 8, cudaMemcpyDeviceToHost));
    if (var_out) VNB_CUDA_OK(cudaMemcpy(var_out, var.p, c * 8, cudaMemcpyDeviceToHost));
  });
}

/* backward of vnb_op_bn_fwd: dz, dgamma, dbeta, dalpha (dalpha may be null when alpha is) from dL/dy */
int vnb_op_bn_bwd(int device, const float* z, const float* dy, const float* gamma, const float* beta, const float* alpha,
                  float* dz, float* dgamma, float* dbeta, float* dalpha, long long voxels, int c) {
  return guarded([&] {
    using namespace vnb;
    need(z, "z"); need(dy, "dy"); need(gamma, "gamma"); need(beta, "beta"); need(dz, "dz");
    if (c % 4 || 256 % (c / 4)) throw std::invalid_argument("vnb_op_bn_bwd: C must be a multiple of 4 with C/4 dividing 256");
    op_device(device);
    const size_t nb = static_cast<size_t>(voxels) * c * 4;
    HostToDev dzb(z, nb), dd(dy, nb), dg(gamma, c * 4), dbt(beta, c * 4), da(alpha, c * 4), dmm(nullptr, c * 4), dmv(nullptr, c * 4);
    DevBuf partial(static_cast<size_t>(kMaxRedBlocks) * 3 * c * 8), mean(c * 8), var(c * 8), scale(c * 4), shift(c * 4);
    DevBuf P(c * 4), Q(c * 4), S(c * 4), gg(c * 4), gb(c * 4), ga(c * 4);
    const unsigned total4 = static_cast<unsigned>(voxels * c / 4);
    const int nblk = op_blocks(total4, 256 * 8, kMaxRedBlocks / 2);
    BnParams bp{};
    bp.gamma[0] = dg.buf.as<float>();
    bp.beta[0] = dbt.buf.as<float>();
    bp.moving_mean[0] = dmm.buf.as<float>();
    bp.moving_var[0] = dmv.buf.as<float>();
    VNB_LAUNCH(bn_stats_v4_kernel, nblk, 256, 0, 0, (const float*)dzb.buf.as<float>(), c, total4, partial.as<double>());
    VNB_LAUNCH(bn_finalize_fwd_kernel, c, 128, 0, 0, (const double*)partial.as<double>(), nblk, 2, c, static_cast<double>(voxels), (int)CH_S, bp,
               0, 0, mean.as<double>(), var.as<double>(), scale.as<float>(), shift.as<float>(), 0);
    BwdArgs b{};
    b.z = dzb.buf.as<float>();
    b.d = dd.buf.as<float>();
    b.scale = scale.as<float>();
    b.shift = shift.as<float>();
    b.alpha = alpha ? da.buf.as<float>() : nullptr;
    b.mean = mean.as<double>();
    b.P = P.as<float>();
    b.Q = Q.as<float>();
    b.S = S.as<float>();
    b.C = c;
    VNB_LAUNCH(bn_bwd_reduce_v4_kernel, nblk, 256, 0, 0, b, total4, partial.as<double>());
    BnGradPtrs gp{};
    gp.dgamma[0] = gg.as<float>();
    gp.dbeta[0] = gb.as<float>();
    gp.dalpha = alpha ? ga.as<float>() : nullptr;
    VNB_LAUNCH(bn_finalize_bwd_kernel, c, 128, 0, 0, (const double*)partial.as<double>(), nblk, c, static_cast<double>(voxels), (int)CH_S, bp,
               (const double*)var.as<double>(), gp, P.as<float>(), Q.as<float>(), S.as<float>(), 0, (const double*)nullptr, 0.0);
    VNB_LAUNCH(bn_bwd_apply_v4_kernel, op_blocks(voxels * c / 4, 256, 2 * kMaxRedBlocks), 256, 0, 0, b, voxels * c);
    op_sync_and_check();
    VNB_CUDA_OK(cudaMemcpy(dz, dd.buf.p, nb, cudaMemcpyDeviceToHost));
    if (dgamma) VNB_CUDA_OK(cudaMemcpy(dgamma, gg.p, c * 4, cudaMemcpyDeviceToHost));
    if (dbeta) VNB_CUDA_OK(cudaMemcpy(dbeta, gb.p, c * 4, cudaMemcpyDeviceToHost));
    if (dalpha && alpha) VNB_CUDA_OK(cudaMemcpy(dalpha, ga.p, c * 4, cudaMemcpyDeviceToHost));
  });
}

/* softmax + one-hot + Dice / Jaccard / cross-entropy loss of model.py:26-92,447,477,495-560 and argmax (model.py:568) on
 * logits [n][voxels][k], labels int32 [n][voxels]; `loss` is a VNB_LOSS_* code.  Outputs optional except loss_out. */
int vnb_op_softmax_dice_fwd(int device, const float* logits, const int32_t* labels, int n, long long voxels, int k, int loss,
                            const float* weights, float alpha, float* loss_out, float* softmax_out, long long* argmax_out,
                            double* terms_out /* [n][k][4] = (I, L, R, X) */) {
  return guarded([&] {
    using namespace vnb;
    need(logits, "logits"); need(labels, "labels"); need(loss_out, "loss_out");
    if (k < 1 || k > kMaxClasses) throw std::invalid_argument("num_classes out of range (1..8)");
    op_device(device);
    EngineConfig ec;
    ec.num_classes = k;
    ec.loss = loss;
    ec.loss_alpha = alpha;
    for (int i = 0; i < kMaxClasses; ++i) ec.loss_weights[i] = (weights && i < k) ? weights[i] : 1.0f;
    const LossCfg lc = loss_cfg_of(ec);
    const size_t nl = static_cast<size_t>(n) * voxels * k * 4;
    HostToDev dl(logits, nl), dlab(labels, static_cast<size_t>(n) * voxels * 4);
    const int nblk = op_blocks(voxels, 2048, 296);
    DevBuf sm(nl), am(static_cast<size_t>(n) * voxels * 8), partial(static_cast<size_t>(n) * nblk * kMaxClasses * 4 * 8);
    DevBuf terms(static_cast<size_t>(n) * kMaxClasses * 4 * 8), coef(static_cast<size_t>(n) * kMaxClasses * 3 * 4), lossd(16);
    dim3 grid(nblk, n);
    VNB_LAUNCH(softmax_loss_fwd_kernel, grid, 256, 0, 0, (const float*)dl.buf.as<float>(), (const int32_t*)dlab.buf.as<int32_t>(), voxels, lc,
               softmax_out ? sm.as<float>() : (float*)nullptr, argmax_out ? am.as<long long>() : (long long*)nullptr, partial.as<double>());
    VNB_LAUNCH(loss_finalize_kernel, 1, 64, 0, 0, (const double*)partial.as<double>(), n, nblk, voxels, lc, terms.as<double>(), coef.as<float>(),
               lossd.as<float>(), (const double*)nullptr, 0, 1.0);
    op_sync_and_check();
    VNB_CUDA_OK(cudaMemcpy(loss_out, lossd.p, 4, cudaMemcpyDeviceToHost));
    if (softmax_out) VNB_CUDA_OK(cudaMemcpy(softmax_out, sm.p, nl, cudaMemcpyDeviceToHost));
    if (argmax_out) VNB_CUDA_OK(cudaMemcpy(argmax_out, am.p, static_cast<size_t>(n) * voxels * 8, cudaMemcpyDeviceToHost));
    if (terms_out) VNB_CUDA_OK(cudaMemcpy(terms_out, terms.p, static_cast<size_t>(n) * k * 4 * 8, cudaMemcpyDeviceToHost));
  });
}

/* dL/dlogits of the same loss */
int vnb_op_softmax_dice_bwd(int device, const float* logits, const int32_t* labels, int n, long long voxels, int k, int loss,
                            const float* weights, float alpha, float* dlogits) {
  return guarded([&] {
    using namespace vnb;
    need(logits, "logits"); need(labels, "labels"); need(dlogits, "dlogits");
    if (k < 1 || k > kMaxClasses) throw std::invalid_argument("num_classes out of range (1..8)");
    op_device(device);
    EngineConfig ec;
    ec.num_classes = k;
    ec.loss = loss;
    ec.loss_alpha = alpha;
    for (int i = 0; i < kMaxClasses; ++i) ec.loss_weights[i] = (weights && i < k) ? weights[i] : 1.0f;
    const LossCfg lc = loss_cfg_of(ec);
    const size_t nl = static_cast<size_t>(n) * voxels * k * 4;
    HostToDev dl(logits, nl), dlab(labels, static_cast<size_t>(n) * voxels * 4);
    const int nblk = op_blocks(voxels, 2048, 296);
    DevBuf partial(static_cast<size_t>(n) * nblk * kMaxClasses * 4 * 8), terms(static_cast<size_t>(n) * kMaxClasses * 4 * 8);
    DevBuf coef(static_cast<size_t>(n) * kMaxClasses * 3 * 4), lossd(16), dg(nl);
    dim3 grid(nblk, n);
    VNB_LAUNCH(softmax_loss_fwd_kernel, grid, 256, 0, 0, (const float*)dl.buf.as<float>(), (const int32_t*)dlab.buf.as<int32_t>(), voxels, lc,
               (float*)nullptr, (long long*)nullptr, partial.as<double>());
    VNB_LAUNCH(loss_finalize_kernel, 1, 64, 0, 0, (const double*)partial.as<double>(), n, nblk, voxels, lc, terms.as<double>(), coef.as<float>(),
               lossd.as<float>(), (const double*)nullptr, 0, 1.0);
    VNB_LAUNCH(softmax_loss_bwd_kernel, grid, 256, 0, 0, (const float*)dl.buf.as<float>(), (const int32_t*)dlab.buf.as<int32_t>(), voxels, lc,
               (const float*)coef.as<float>(), 1.0f, dg.as<float>());
    op_sync_and_check();
    VNB_CUDA_OK(cudaMemcpy(dlogits, dg.p, nl, cudaMemcpyDeviceToHost));
  });
}

/* one tf.train.AdamOptimizer step (epsilon-hat form, model.py:652) on a flat parameter vector, in place:
 * t = 1-based step count, lr = the decayed learning rate of that step */
int vnb_op_adam(int device, float* p, const float* g, float* m, float* v, long long count, float lr, long long t) {
  return guarded([&] {
    using namespace vnb;
    need(p, "p"); need(g, "g"); need(m, "m"); need(v, "v");
    op_device(device);
    const size_t nb = static_cast<size_t>(count) * 4;
    HostToDev dp(p, nb), dg(g, nb), dm(m, nb), dv(v, nb);
    const double b1 = 0.9, b2 = 0.999;
    const float lr_t = static_cast<float>(lr * std::sqrt(1.0 - std::pow(b2, (double)t)) / (1.0 - std::pow(b1, (double)t)));
    VNB_LAUNCH(adam_step_kernel, op_blocks(count, 256, 2 * kMaxRedBlocks), 256, 0, 0, dp.buf.as<float>(), (const float*)dg.buf.as<float>(),
               dm.buf.as<float>(), dv.buf.as<float>(), count, lr_t, 0.9f, 0.999f, 1e-8f, 1.0f);
    op_sync_and_check();
    VNB_CUDA_OK(cudaMemcpy(p, dp.buf.p, nb, cudaMemcpyDeviceToHost));
    VNB_CUDA_OK(cudaMemcpy(m, dm.buf.p, nb, cudaMemcpyDeviceToHost));
    VNB_CUDA_OK(cudaMemcpy(v, dv.buf.p, nb, cudaMemcpyDeviceToHost));
  });
}

}  // extern "C"
