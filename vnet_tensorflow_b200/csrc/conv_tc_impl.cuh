// tcgen05 convolution path: engine glue (plans, weight packing, launches) and standalone op hooks.
#pragma once
#include "conv_tc.cuh"
#include "engine.cuh"

namespace vnb {

inline int tc_query_sms() {
#ifdef VNB_EMULATE
  return 2;  // emulation: two persistent "CTAs" so the tile scheduler's striding is exercised
#else
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
#endif
}

// fill the tensor maps of a plan; lo pointers may be null (single-pass bf16): hi is reused as a dummy
inline void tc_encode_plan(TcKernelPlan& pl, int Nmax, const uint16_t* a1_hi, const uint16_t* a1_lo, const uint16_t* a2_hi,
                           const uint16_t* a2_lo) {
  const TcGeom& g = pl.g;
  const int ks = pl.KS;
  const int lines = g.resident ? g.bh + ks - 1 : g.bh;   // resident mode: the box carries the halo lines
  tma_encode_act(&pl.a1_hi, a1_hi, Nmax, g.D, g.H, g.W, g.C1, pl.KC, lines, g.bd, g.LP);
  tma_encode_act(&pl.a1_lo, a1_lo ? a1_lo : a1_hi, Nmax, g.D, g.H, g.W, g.C1, pl.KC, lines, g.bd, g.LP);
  if (g.C2 > 0) {
    tma_encode_act(&pl.a2_hi, a2_hi, Nmax, g.D, g.H, g.W, g.C2, pl.KC, lines, g.bd, g.LP);
    tma_encode_act(&pl.a2_lo, a2_lo ? a2_lo : a2_hi, Nmax, g.D, g.H, g.W, g.C2, pl.KC, lines, g.bd, g.LP);
  } else {
    pl.a2_hi = pl.a1_hi;
    pl.a2_lo = pl.a1_lo;
  }
  const long long rows = static_cast<long long>(g.n_slices) * ks * ks * g.n_kc * ks * pl.CT;
  tma_encode_w(&pl.w_hi, pl.wp_hi, rows, pl.KC, ks * pl.CT);
  tma_encode_w(&pl.w_lo, pl.wp_lo ? pl.wp_lo : pl.wp_hi, rows, pl.KC, ks * pl.CT);
}

inline void Engine::tc_setup() {
  sm_count_ = tc_query_sms();
  const bool lo = cfg_.precision == PREC_BF16X3;
  const int NB = cfg_.max_batch;
  size_t wg_partial_floats = 0;
  for (Unit& u : units_) {
    if (u.kind != U_CONV5 && u.kind != U_CONV3) continue;
    const int ks = u.kind == U_CONV3 ? 3 : 5;
    const Act& x1 = acts_[u.in1];
    const Act& o = acts_[u.out];
    const Dims d = o.dims;
    // multi-modal network input (networks.py:260-266): the image's bf16 copies are zero-padded to 16 channels
    const bool padded_in = (u.in1 == image_act_) && image_cpad_ > 0;
    const int c1 = padded_in ? image_cpad_ : u.Cin1;
    TcKernelPlan& f = u.tc.fprop;
    if (tc_plan_geometry(f, NB, d.D, d.H, d.W, c1, u.Cin2, u.Cout, 0, lo, ks, sm_count_)) {
      f.wp_elems = static_cast<size_t>(ks * ks * ks) * (c1 + u.Cin2) * u.Cout;
      f.wp_hi = dev_alloc<uint16_t>(f.wp_elems);
      f.wp_lo = lo ? dev_alloc<uint16_t>(f.wp_elems) : nullptr;
      tc_encode_plan(f, NB, x1.a_hi, x1.a_lo, u.in2 >= 0 ? acts_[u.in2].a_hi : nullptr, u.in2 >= 0 ? acts_[u.in2].a_lo : nullptr);
      f.valid = true;
    }
    TcKernelPlan& g = u.tc.dgrad;
    if (u.need_dgrad && tc_plan_geometry(g, NB, d.D, d.H, d.W, u.Cout, 0, u.Cin1, u.Cin2, lo, ks, sm_count_)) {
      g.wp_elems = u.w_count;
      g.wp_hi = dev_alloc<uint16_t>(g.wp_elems);
      g.wp_lo = lo ? dev_alloc<uint16_t>(g.wp_elems) : nullptr;
      tc_encode_plan(g, NB, o.d_hi, o.d_lo, nullptr, nullptr);
      g.valid = true;
    }
    WgPlan& wg = u.tc.wgrad;
    if (wg_plan_geometry(wg, NB, d.D, d.H, d.W, c1, u.Cin2, u.Cout, lo, sm_count_, ks)) {
      wg_encode_plan(wg, NB, x1.a_hi, x1.a_lo, u.in2 >= 0 ? acts_[u.in2].a_hi : nullptr,
                     u.in2 >= 0 ? acts_[u.in2].a_lo : nullptr, o.d_hi, o.d_lo);
      wg.valid = true;
      wg_partial_floats = std::max(wg_partial_floats, wg.partial_floats);
    }
    WdPlan& wd = u.tc.wgrad_deep;
    if (ks == 5 && !padded_in && wd_plan_geometry(wd, NB, d.D, d.H, d.W, c1, u.Cin2, u.Cout, lo, sm_count_)) {
      wd_encode_plan(wd, NB, x1.a_hi, x1.a_lo, u.in2 >= 0 ? acts_[u.in2].a_hi : nullptr, u.in2 >= 0 ? acts_[u.in2].a_lo : nullptr,
                     o.d_hi, o.d_lo);
      wd.valid = true;
      wg_partial_floats = std::max(wg_partial_floats, wd.partial_floats);
    }
  }
  if (wg_partial_floats) wg_partial_ = dev_alloc<float>(wg_partial_floats);
  // 2^3 stride-2 units: tcgen05 / TMA gather and scatter plans (k2_tc.cuh) over the fp32 activations and gradients
  if (k2tc_enabled()) {
    std::vector<K2PackJob> jobs;
    for (Unit& u : units_) {
      if (u.kind != U_DOWN && u.kind != U_UP) continue;
      const Act& x1 = acts_[u.in1];
      const Act& o = acts_[u.out];
      const bool down = u.kind == U_DOWN;
      const int CF = down ? u.Cin1 : u.Cout, CC = down ? u.Cout : u.Cin1;
      const Dims cd = down ? o.dims : x1.dims;
      K2PackJob j{};
      j.w = params_ + u.w_off;
      j.CF = CF;
      j.CC = CC;
      const size_t img_elems = 3 * u.w_count;
      // fprop: down = gather (fine x -> coarse z), up = scatter (coarse x -> fine z)
      if (k2tc_plan_geometry(u.k2_fprop, !down, NB, cd, CF, CC)) {
        u.k2_fprop.img = dev_alloc<uint16_t>(img_elems);
        (down ? j.img_gather : j.img_scatter) = u.k2_fprop.img;
        k2tc_encode_plan(u.k2_fprop, NB, cd, CF, CC, down ? x1.a : u.z, down ? u.z : x1.a);
        u.k2_fprop.valid = true;
      }
      // dgrad: down = scatter (coarse dz -> fine dx), up = gather (fine dz -> coarse dx)
      if (u.need_dgrad && k2tc_plan_geometry(u.k2_dgrad, down, NB, cd, CF, CC)) {
        u.k2_dgrad.img = dev_alloc<uint16_t>(img_elems);
        (down ? j.img_scatter : j.img_gather) = u.k2_dgrad.img;
        k2tc_encode_plan(u.k2_dgrad, NB, cd, CF, CC, down ? x1.d : o.d, down ? o.d : x1.d);
        u.k2_dgrad.valid = true;
      }
      // filter gradient: fine = x (down) or dz (up), coarse = dz (down) or x (up)
      if (k2wg_plan_geometry(u.k2_wgrad, NB, cd, CF, CC)) {
        k2wg_encode_plan(u.k2_wgrad, NB, cd, CF, CC, down ? x1.a : o.d, down ? o.d : x1.a);
        u.k2_wgrad.valid = true;
      }
      if (!j.img_gather && !j.img_scatter) continue;
      j.first_block = k2_pack_blocks_;
      j.n_blocks = static_cast<int>(std::min<size_t>((u.w_count + 255) / 256, 64));
      k2_pack_blocks_ += j.n_blocks;
      jobs.push_back(j);
    }
    if (!jobs.empty()) {
      k2_pack_njobs_ = static_cast<int>(jobs.size());
      k2_pack_jobs_dev_ = dev_alloc<K2PackJob>(jobs.size());
      VNB_CUDA_OK(cudaMemcpy(k2_pack_jobs_dev_, jobs.data(), jobs.size() * sizeof(K2PackJob), cudaMemcpyHostToDevice));
    }
  }
}

// Weight packing (fp32 master -> bf16 hi/lo GEMM-B tiles) after every optimiser step.  The packs of the first forward
// units (a few percent of the bytes) run on the compute stream; everything else -- the deep layers' forward packs and all
// input-gradient packs, ~95 % of the 0.7 GB this pass moves -- runs on the filter-gradient side stream (idle during the
// forward pass) under the tensor-bound first convolutions, and the compute stream waits for it right before the first
// unit that needs it (Engine::tc_wait_late_packs).
inline void Engine::tc_prepare_weights() {
  if (!weights_dirty_) return;
  if (!pack_built_) {  // job tables, built once
    std::vector<PackJob> early, late;
    int eb = 0, lb = 0;
    size_t seen = 0, total = 0;
    for (const Unit& u : units_)
      if (u.kind == U_CONV5 || u.kind == U_CONV3) total += u.w_count;
    pack_first_late_unit_ = static_cast<int>(units_.size());
    for (size_t ui = 0; ui < units_.size(); ++ui) {
      Unit& u = units_[ui];
      if (u.kind != U_CONV5 && u.kind != U_CONV3) continue;
      const bool is_early = wg_stream_ && (seen + u.w_count) * 20 <= total;   // first <= 5 % of the filter elements
      seen += u.w_count;
      if (!is_early && pack_first_late_unit_ == static_cast<int>(units_.size())) pack_first_late_unit_ = static_cast<int>(ui);
      for (int pass = 0; pass < 2; ++pass) {
        TcKernelPlan& pl = pass == 0 ? u.tc.fprop : u.tc.dgrad;
        if (!pl.valid) continue;
        const bool e = is_early && pass == 0 && pack_first_late_unit_ == static_cast<int>(units_.size());
        std::vector<PackJob>& jobs = e ? early : late;
        int& blocks = e ? eb : lb;
        PackJob j;
        j.w = params_ + u.w_off;
        j.hi = pl.wp_hi;
        j.lo = pl.wp_lo;
        j.Cin = u.Cin1 + u.Cin2;
        j.Cout = u.Cout;
        j.dgrad = pass;
        j.CT = pl.CT;
        j.KC = pl.KC;
        j.Cin_gemm = pl.g.C1 + pl.g.C2;
        j.first_block = blocks;
        j.KS = pl.KS;
        blocks += pl.g.n_slices * pl.KS * pl.KS * pl.g.n_kc;
        jobs.push_back(j);
      }
    }
    pack_blocks_[0] = eb;
    pack_blocks_[1] = lb;
    pack_njobs_[0] = static_cast<int>(early.size());
    pack_njobs_[1] = static_cast<int>(late.size());
    for (int k = 0; k < 2; ++k) {
      const std::vector<PackJob>& jobs = k == 0 ? early : late;
      if (jobs.empty()) continue;
      pack_jobs_dev_[k] = dev_alloc<PackJob>(jobs.size());
      VNB_CUDA_OK(cudaMemcpy(pack_jobs_dev_[k], jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice));
    }
    pack_built_ = true;
  }
  if (pack_njobs_[0]) {
    VNB_LAUNCH(pack_w5_multi_kernel, pack_blocks_[0], 256, 0, stream_, (const PackJob*)pack_jobs_dev_[0], pack_njobs_[0]);
    ++launches_;
  }
  if (k2_pack_njobs_) {   // 2^3 filters (0.2 % of the parameters): one small launch on the compute stream
    VNB_LAUNCH(k2tc_pack_multi_kernel, k2_pack_blocks_, 256, 0, stream_, (const K2PackJob*)k2_pack_jobs_dev_, k2_pack_njobs_);
    ++launches_;
  }
  if (pack_njobs_[1]) {
    cudaStream_t st = stream_;
#ifndef VNB_EMULATE
    if (wg_stream_ && pack_njobs_[0]) {   // after the optimiser step (compute stream), beside the first forward units
      VNB_CUDA_OK(cudaEventRecord(wg_ready_ev_, stream_));
      VNB_CUDA_OK(cudaStreamWaitEvent(wg_stream_, wg_ready_ev_, 0));
      st = wg_stream_;
    }
#endif
    VNB_LAUNCH(pack_w5_multi_kernel, pack_blocks_[1], 256, 0, st, (const PackJob*)pack_jobs_dev_[1], pack_njobs_[1]);
    ++launches_;
#ifndef VNB_EMULATE
    if (st != stream_) {
      VNB_CUDA_OK(cudaEventRecord(pack_done_ev_, st));
      late_packs_pending_ = true;
    }
#endif
  }
  weights_dirty_ = false;
}

// called by forward() before unit `ui` (and by the backward pass): the compute stream joins the side-stream pack
inline void Engine::tc_wait_late_packs(int ui) {
#ifndef VNB_EMULATE
  if (late_packs_pending_ && ui >= pack_first_late_unit_) {
    VNB_CUDA_OK(cudaStreamWaitEvent(stream_, pack_done_ev_, 0));
    late_packs_pending_ = false;
  }
#else
  (void)ui;
#endif
}

inline void Engine::tc_run_fprop(Unit& u, int N) {
  TcKernelPlan& pl = u.tc.fprop;
  TcArgs a;
  a.g = pl.g;
  a.g.N = N;
  a.g.n_items = N * a.g.n_db * a.g.n_hb * a.g.n_wb * a.g.n_slices;
  a.bias = params_ + u.b_off;
  a.res = u.res >= 0 ? acts_[u.res].a : nullptr;
  a.out1 = u.z;
  a.out2 = nullptr;
  a.acc1 = a.acc2 = 0;
  // column kernel: the epilogue also produces the per-CTA (sum z, sum z^2) rows of the unit's batch norm
  fused_stats_blocks_ = 0;
  if (pl.col && !u.bn_inference && !getenv("VNB_NO_FUSED_STATS")) {
    a.stats = partial_;
    fused_stats_blocks_ = std::max(1, std::min(N * pl.cg.n_hb * pl.cg.n_seg, sm_count_));
  }
  ProfScope ps(*this, 0, conv5_flops(u, N), 0, &u, "fprop");
  launches_ += tc_launch(pl, a, cfg_.precision == PREC_BF16X3, sm_count_, stream_);
}

inline void Engine::tc_run_dgrad(Unit& u, int N) {
  TcKernelPlan& pl = u.tc.dgrad;
  TcArgs a;
  a.g = pl.g;
  a.g.N = N;
  a.g.n_items = N * a.g.n_db * a.g.n_hb * a.g.n_wb * a.g.n_slices;
  a.bias = nullptr;
  a.res = nullptr;
  a.out1 = acts_[u.in1].d;
  a.acc1 = u.in1_accumulate ? 1 : 0;
  a.out2 = u.in2 >= 0 ? acts_[u.in2].d : nullptr;
  a.acc2 = u.in2_accumulate ? 1 : 0;
  ProfScope ps(*this, 0, conv5_flops(u, N), 0, &u, "dgrad");
  launches_ += tc_launch(pl, a, cfg_.precision == PREC_BF16X3, sm_count_, stream_);
}

inline void Engine::tc_run_wgrad(Unit& u, int N) {
  cudaStream_t st = wgrad_stream_begin();
  ProfScope ps(*this, 1, conv5_flops(u, N), st, &u, "wgrad");
  if (u.tc.wgrad_deep.valid) {
    launches_ += wd_launch(u.tc.wgrad_deep, N, wg_partial_, grads_ + u.w_off, sm_count_, st);
    return;
  }
  wg_launch(u.tc.wgrad, N, cfg_.precision == PREC_BF16X3, wg_partial_, grads_ + u.w_off, st, u.Cin1 + u.Cin2);
  launches_ += 2;
}

// ---- standalone op hooks (tests): fp32 device buffers in, fp32 out ---------------------------------
struct TcScratch {
  std::vector<void*> ptrs;
  template <class T>
  T* alloc(size_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(n, 8) * sizeof(T)) != cudaSuccess) throw std::runtime_error("CUDA: out of memory in tc op hook");
    ptrs.push_back(p);
    return static_cast<T*>(p);
  }
  ~TcScratch() {
    for (void* p : ptrs) cudaFree(p);
  }
};

inline void tc_op_conv5(int precision, const float* x, const float* w, const float* bias, const float* res, float* y, int n,
                        Dims dims, int cin, int cout, bool dgrad_form, int ks = 5) {
  const bool lo = precision == PREC_BF16X3;
  const int ci = dgrad_form ? cout : cin, co = dgrad_form ? cin : cout;
  TcKernelPlan pl;
  if (!tc_plan_geometry(pl, n, dims.D, dims.H, dims.W, ci, 0, co, 0, lo, ks, tc_query_sms()))
    throw std::invalid_argument("shape not supported by the tensor-core convolution (channels % 16, line blocking)");
  TcScratch s;
  const size_t nx = static_cast<size_t>(n) * dims.D * dims.H * dims.W * ci;
  uint16_t* xh = s.alloc<uint16_t>(nx);
  uint16_t* xl = lo ? s.alloc<uint16_t>(nx) : nullptr;
  VNB_LAUNCH(split_bf16_kernel, 1024, 256, 0, 0, x, static_cast<long long>(nx), xh, xl);
  pl.wp_elems = static_cast<size_t>(ks * ks * ks) * cin * cout;
  pl.wp_hi = s.alloc<uint16_t>(pl.wp_elems);
  pl.wp_lo = lo ? s.alloc<uint16_t>(pl.wp_elems) : nullptr;
  VNB_LAUNCH(pack_w5_kernel, pl.g.n_slices * ks * ks * pl.g.n_kc, 256, 0, 0, w, cin, cout, dgrad_form ? 1 : 0, pl.CT, pl.KC, pl.wp_hi,
             pl.wp_lo, cin, ks);
  tc_encode_plan(pl, n, xh, xl, nullptr, nullptr);
  TcArgs a;
  a.g = pl.g;
  a.bias = bias;
  a.res = res;
  a.out1 = y;
  a.out2 = nullptr;
  a.acc1 = a.acc2 = 0;
  tc_launch(pl, a, lo, tc_query_sms(), 0);
  if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess)
    throw std::runtime_error("CUDA: tensor-core convolution kernel failed");
}

// 2^3 stride-2 gather / scatter through the tcgen05 kernel; false when the shape is outside its domain
inline bool tc_op_k2(bool scatter, float* fine, float* coarse, const float* w, const float* bias, int n, Dims cd, int cf, int cc,
                     bool accumulate = false) {
  K2TcPlan pl;
  if (!k2tc_enabled() || !k2tc_plan_geometry(pl, scatter, n, cd, cf, cc)) return false;
  TcScratch s;
  const size_t wn = static_cast<size_t>(8) * cf * cc;
  pl.img = s.alloc<uint16_t>(3 * wn);
  K2PackJob j{};
  j.w = w;
  (scatter ? j.img_scatter : j.img_gather) = pl.img;
  j.CF = cf;
  j.CC = cc;
  j.first_block = 0;
  j.n_blocks = static_cast<int>(std::min<size_t>((wn + 255) / 256, 64));
  K2PackJob* jd = s.alloc<K2PackJob>(1);
  if (cudaMemcpy(jd, &j, sizeof(j), cudaMemcpyHostToDevice) != cudaSuccess) throw std::runtime_error("CUDA: copy failed in tc_op_k2");
  VNB_LAUNCH(k2tc_pack_multi_kernel, j.n_blocks, 256, 0, 0, (const K2PackJob*)jd, 1);
  k2tc_encode_plan(pl, n, cd, cf, cc, fine, coarse);
  k2tc_launch(pl, n, bias, accumulate, tc_query_sms(), 0);
  if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess)
    throw std::runtime_error("CUDA: tensor-core 2x2x2 kernel failed");
  return true;
}

// 2^3 filter gradient through the tcgen05 kernel (dw is overwritten); false when the shape is outside its domain
inline bool tc_op_k2_wgrad(const float* fine, const float* coarse, float* dw, int n, Dims cd, int cf, int cc) {
  K2WgPlan pl;
  if (!k2tc_enabled() || !k2wg_plan_geometry(pl, n, cd, cf, cc)) return false;
  if (cudaMemset(dw, 0, static_cast<size_t>(8) * cf * cc * 4) != cudaSuccess) throw std::runtime_error("CUDA: memset failed in tc_op_k2_wgrad");
  k2wg_encode_plan(pl, n, cd, cf, cc, fine, coarse);
  k2wg_launch(pl, n, dw, tc_query_sms(), 0);
  if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess)
    throw std::runtime_error("CUDA: tensor-core 2x2x2 filter-gradient kernel failed");
  return true;
}

inline void tc_op_wgrad5(int precision, const float* x, const float* dy, float* dw, int n, Dims dims, int cin, int cout, int ks = 5) {
  const bool lo = precision == PREC_BF16X3;
  WgPlan pl;
  WdPlan wd;
  const bool deep = ks == 5 && wd_plan_geometry(wd, n, dims.D, dims.H, dims.W, cin, 0, cout, lo, tc_query_sms());
  if (!deep && !wg_plan_geometry(pl, n, dims.D, dims.H, dims.W, cin, 0, cout, lo, tc_query_sms(), ks))
    throw std::invalid_argument("shape not supported by the tensor-core wgrad (channels % 16, W in {8..128})");
  TcScratch s;
  const size_t V = static_cast<size_t>(n) * dims.D * dims.H * dims.W;
  uint16_t* xh = s.alloc<uint16_t>(V * cin);
  uint16_t* xl = lo ? s.alloc<uint16_t>(V * cin) : nullptr;
  uint16_t* zh = s.alloc<uint16_t>(V * cout);
  uint16_t* zl = lo ? s.alloc<uint16_t>(V * cout) : nullptr;
  VNB_LAUNCH(split_bf16_kernel, 1024, 256, 0, 0, x, static_cast<long long>(V * cin), xh, xl);
  VNB_LAUNCH(split_bf16_kernel, 1024, 256, 0, 0, dy, static_cast<long long>(V * cout), zh, zl);
  if (deep) {   // per-tap GEMM form of the deep levels (wgrad_deep.cuh)
    float* partial = s.alloc<float>(wd.partial_floats);
    wd_encode_plan(wd, n, xh, xl, nullptr, nullptr, zh, zl);
    wd_launch(wd, n, partial, dw, tc_query_sms(), 0);
    if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess)
      throw std::runtime_error("CUDA: tensor-core deep-level wgrad kernel failed");
    return;
  }
  float* partial = s.alloc<float>(pl.partial_floats);
  wg_encode_plan(pl, n, xh, xl, nullptr, nullptr, zh, zl);
  wg_launch(pl, n, lo, partial, dw, 0);
  if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess)
    throw std::runtime_error("CUDA: tensor-core wgrad kernel failed");
}

}  // namespace vnb
