// Kernel micro-benchmark for the tensor-core convolution kernels (development tool, not part of the product):
//   kbench wgrad|fprop N D H W Cin Cout precision(1=bf16,2=bf16x3) reps [ks]
// Times `reps` back-to-back launches with CUDA events on random bf16 operands and prints us / launch and the
// algorithmic TFLOP/s (2*ks^3*Cin*Cout*voxels).  VNB_KB_DBG=1 prints the MMA-warp cycle counters of CTA 0.
#define VNB_KB_COUNTERS 1   // MMA-warp cycle counters (compiled out of the product library)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../vnet_tensorflow_b200/csrc/engine.cuh"
#include "../vnet_tensorflow_b200/csrc/conv_tc_impl.cuh"

using namespace vnb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void fill_bf16(uint16_t* p, size_t n, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    p[i] = f32_to_bf16(((h & 0xFFFF) / 65536.0f - 0.5f));
  }
}

int main(int argc, char** argv) {
  if (argc < 10) { printf("usage: kbench wgrad|fprop N D H W Cin Cout precision reps [ks] [cin2]\n"); return 2; }
  const std::string op = argv[1];
  const int N = atoi(argv[2]), D = atoi(argv[3]), H = atoi(argv[4]), W = atoi(argv[5]), Cin = atoi(argv[6]), Cout = atoi(argv[7]);
  const int prec = atoi(argv[8]), reps = atoi(argv[9]);
  const int ks = argc > 10 ? atoi(argv[10]) : 5;
  const int cin2 = argc > 11 ? atoi(argv[11]) : 0;
  const bool lo = prec == 2;
  const size_t V = (size_t)N * D * H * W;
  const int sms = tc_query_sms();
  TcScratch s;
  auto mk = [&](size_t n, uint32_t seed) { uint16_t* p = s.alloc<uint16_t>(n); fill_bf16<<<1024, 256>>>(p, n, seed); return p; };
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double flops = 2.0 * ks * ks * ks * (Cin + cin2) * (double)Cout * (double)V;
  float ms = 0;
  if (op == "k2g" || op == "k2s" || op == "k2w") {
    // 2x2x2 stride-2 kernels: (D, H, W) are the COARSE dims, Cin = fine channels CF, Cout = coarse channels CC
    const size_t Vc = V, Vf = V * 8;
    float* fine = s.alloc<float>(Vf * Cin);
    float* coarse = s.alloc<float>(Vc * Cout);
    float* wt = s.alloc<float>((size_t)8 * Cin * Cout);
    float* dwt = s.alloc<float>((size_t)8 * Cin * Cout);
    CK(cudaMemset(fine, 0, Vf * Cin * 4)); CK(cudaMemset(coarse, 0, Vc * Cout * 4)); CK(cudaMemset(wt, 0, (size_t)8 * Cin * Cout * 4));
    CK(cudaMemset(dwt, 0, (size_t)8 * Cin * Cout * 4));
    K2Args p{};
    p.fine_in = fine; p.coarse_in = coarse; p.fine_out = fine; p.coarse_out = coarse; p.w = wt; p.dw = dwt; p.bias = nullptr;
    p.CF = Cin; p.CC = Cout; p.cd = Dims{D, H, W}; p.N = N; p.accumulate = 0;
    const long long M = (long long)Vc;
    // tcgen05 / TMA form (k2_tc.cuh) unless VNB_K2_NO_TC is set; `cin2` != 0 selects the accumulating (reduce-add) epilogue
    K2TcPlan tcp;
    if (op != "k2w" && k2tc_enabled() && k2tc_plan_geometry(tcp, op == "k2s", N, p.cd, Cin, Cout)) {
      tcp.img = s.alloc<uint16_t>((size_t)24 * Cin * Cout);
      CK(cudaMemset(tcp.img, 0, (size_t)48 * Cin * Cout));
      k2tc_encode_plan(tcp, N, p.cd, Cin, Cout, fine, coarse);
      tcp.valid = true;
      printf("k2tc plan: tile %dx%dx%d items=%d n_kc=%d NB=%d n_nb=%d stages=%d stage=%d smem=%zu tmem=%d\n", tcp.g.ow_t, tcp.g.oh_t, tcp.g.od_t,
             tcp.g.n_items, tcp.g.n_kc, tcp.g.NB, tcp.g.n_nb, tcp.g.stages, tcp.g.stage_bytes, tcp.smem, tcp.g.tmem_cols);
    }
    K2WgPlan wgp;
    if (op == "k2w" && k2tc_enabled() && k2wg_plan_geometry(wgp, N, p.cd, Cin, Cout)) {
      k2wg_encode_plan(wgp, N, p.cd, Cin, Cout, fine, coarse);
      wgp.valid = true;
      printf("k2wg plan: tile %dx%dx%d n_mb=%d sets=%d set=%d smem=%zu tmem=%d\n", wgp.g.ow_t, wgp.g.oh_t, wgp.g.od_t, wgp.g.n_mb * wgp.g.n_cb, wgp.g.sets,
             wgp.g.set_bytes, wgp.smem, wgp.g.tmem_cols);
    }
    auto launch = [&]() {
      if (wgp.valid) {
        k2wg_launch(wgp, N, dwt, sms, 0);
      } else if (tcp.valid) {
        k2tc_launch(tcp, N, nullptr, cin2 != 0, sms, 0);
      } else if (op == "k2g") {
        dim3 grid((unsigned)((M + kK2_BM - 1) / kK2_BM), (p.CC + kK2_BN - 1) / kK2_BN);
        k2_gather_mma_kernel<<<grid, 256>>>(p, M);
      } else if (op == "k2s") {
        dim3 grid((unsigned)((M + kK2_BM - 1) / kK2_BM), (8 * p.CF + kK2_BN - 1) / kK2_BN);
        k2_scatter_mma_kernel<<<grid, 256>>>(p, M);
      } else {
        const int gx = (8 * p.CF + kK2_BM - 1) / kK2_BM, gy = (p.CC + kK2_BN - 1) / kK2_BN;
        long long splits = std::max<long long>(1, std::min<long long>((M + 255) / 256, (4 * 148 + gx * gy - 1) / (gx * gy)));
        long long mps = ((M + splits - 1) / splits + kK2_BK - 1) / kK2_BK * kK2_BK;
        splits = (M + mps - 1) / mps;
        dim3 grid(gx, gy, (unsigned)splits);
        k2_wgrad_mma_kernel<<<grid, 256>>>(p, M, mps);
      }
    };
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us2 = ms * 1e3 / reps;
    const double bytes = (double)Vf * Cin * 4 + (double)Vc * Cout * 4;
    printf("KBENCH %s%s N=%d coarse %dx%dx%d CF=%d CC=%d : %.1f us/launch  %.0f GB/s (fine + coarse tensor once)\n", op.c_str(),
           (tcp.valid || wgp.valid) ? (cin2 ? " [tcgen05, reduce-add]" : " [tcgen05]") : "", N, D, H, W, Cin, Cout, us2, bytes / (us2 * 1e-6) / 1e9);
    return 0;
  } else if (op == "wgrad" && ks == 5 && [&] { WdPlan t; return wd_plan_geometry(t, N, D, H, W, Cin, cin2, Cout, lo, sms); }()) {
    // deep levels: per-tap GEMM kernel (wgrad_deep.cuh); VNB_WG_NO_DEEP=1 times wgrad5_tc_kernel instead
    WdPlan pl;
    wd_plan_geometry(pl, N, D, H, W, Cin, cin2, Cout, lo, sms);
    uint16_t *xh = mk(V * Cin, 1), *xl = lo ? mk(V * Cin, 2) : nullptr, *zh = mk(V * Cout, 3), *zl = lo ? mk(V * Cout, 4) : nullptr;
    uint16_t *x2h = cin2 ? mk(V * cin2, 5) : nullptr, *x2l = (cin2 && lo) ? mk(V * cin2, 6) : nullptr;
    float* partial = s.alloc<float>(pl.partial_floats);
    float* dw = s.alloc<float>((size_t)125 * (Cin + cin2) * Cout);
    wd_encode_plan(pl, N, xh, xl, x2h, x2l, zh, zl);
    printf("wgrad deep plan: HT=%d n_hb=%d n_cib=%d n_cob=%d NB=%d ksplit=%d stages=%d stage=%d items=%d smem=%zu\n", pl.g.HT, pl.g.n_hb, pl.g.n_cib,
           pl.g.n_cob, pl.g.NB, pl.g.ksplit, pl.g.stages, pl.g.stage_bytes, pl.g.n_items, pl.smem);
    for (int i = 0; i < 3; ++i) wd_launch(pl, N, partial, dw, sms, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) wd_launch(pl, N, partial, dw, sms, 0);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  } else if (op == "wgrad") {
    WgPlan pl;
    if (!wg_plan_geometry(pl, N, D, H, W, Cin, cin2, Cout, lo, sms, ks)) { printf("unsupported shape\n"); return 1; }
    uint16_t *xh = mk(V * Cin, 1), *xl = lo ? mk(V * Cin, 2) : nullptr, *zh = mk(V * Cout, 3), *zl = lo ? mk(V * Cout, 4) : nullptr;
    uint16_t *x2h = cin2 ? mk(V * cin2, 5) : nullptr, *x2l = (cin2 && lo) ? mk(V * cin2, 6) : nullptr;
    float* partial = s.alloc<float>(pl.partial_floats);
    float* dw = s.alloc<float>((size_t)ks * ks * ks * (Cin + cin2) * Cout);
    wg_encode_plan(pl, N, xh, xl, x2h, x2l, zh, zl);
    long long* dbg = nullptr;
    if (getenv("VNB_KB_DBG")) { dbg = s.alloc<long long>(8 * 1024); CK(cudaMemset(dbg, 0, 8 * 1024 * sizeof(long long))); pl.g.dbg = dbg; }
    printf("wgrad plan: nch=%d swap=%d q_split=%d Wr=%d n_wb=%d HT=%d n_hb=%d z_stages=%d groups=%d splits=%d+%d pairs=%d grid=%d smem=%zu xt=%d zt=%d\n",
           pl.g.nch, pl.g.swap, pl.g.q_split, pl.g.Wr, pl.g.n_wb, pl.g.HT, pl.g.n_hb, pl.g.z_stages, pl.g.ngroups, pl.g.splits[0], pl.g.splits[1],
           pl.g.n_pc * pl.g.n_qg, (pl.g.splits[0] + pl.g.splits[1]) * pl.g.n_pc * pl.g.n_qg, pl.smem, pl.g.xt_bytes, pl.g.zt_bytes);
    for (int i = 0; i < 3; ++i) wg_launch(pl, N, lo, partial, dw, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) wg_launch(pl, N, lo, partial, dw, 0);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (dbg) {
      std::vector<long long> h(8 * 1024);
      CK(cudaMemcpy(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      for (int c : {0, 1, sms / 2, sms - 1}) {
        const long long* d = &h[8 * c];
        printf("  cta %3d: loop cycles %lld  wait_full cycles %lld  mmas %lld  -> %.1f cyc/mma, wall %.1f us => %.0f MHz\n", c, d[0], d[1], d[2],
               d[2] ? (double)d[0] / d[2] : 0.0, d[3] / 1e3, d[3] ? d[0] / (d[3] / 1e3) : 0.0);
      }
    }
  } else {
    TcKernelPlan pl;
    if (!tc_plan_geometry(pl, N, D, H, W, Cin, cin2, Cout, 0, lo, ks, sms)) { printf("unsupported shape\n"); return 1; }
    uint16_t *xh = mk(V * Cin, 1), *xl = lo ? mk(V * Cin, 2) : nullptr;
    uint16_t *x2h = cin2 ? mk(V * cin2, 5) : nullptr, *x2l = (cin2 && lo) ? mk(V * cin2, 6) : nullptr;
    pl.wp_elems = (size_t)ks * ks * ks * (Cin + cin2) * Cout;
    pl.wp_hi = mk(pl.wp_elems, 7);
    pl.wp_lo = lo ? mk(pl.wp_elems, 8) : nullptr;
    float* y = s.alloc<float>(V * Cout);
    tc_encode_plan(pl, N, xh, xl, x2h, x2l);
    TcArgs a;
    a.g = pl.g; a.bias = nullptr; a.res = nullptr; a.out1 = y; a.out2 = nullptr; a.acc1 = a.acc2 = 0;
    long long* dbg = nullptr;
    if (getenv("VNB_KB_DBG")) { dbg = s.alloc<long long>(8 * 1024); CK(cudaMemset(dbg, 0, 8 * 1024 * sizeof(long long))); a.dbg = dbg; }
    printf("fprop plan: col=%d (ds=%d n_seg=%d) CT=%d KC=%d T=%d bh=%d bd=%d LP=%d lpt=%d resident=%d n_a=%d n_b=%d items=%d smem=%zu\n", (int)pl.col,
           pl.cg.ds, pl.cg.n_seg, pl.CT, pl.KC, pl.g.T, pl.g.bh, pl.g.bd, pl.g.LP, pl.g.lpt, pl.g.resident, pl.g.n_a, pl.g.n_b, pl.g.n_items, pl.smem);
    for (int i = 0; i < 3; ++i) tc_launch(pl, a, lo, sms, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) tc_launch(pl, a, lo, sms, 0);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (dbg) {
      std::vector<long long> h(8 * 1024);
      CK(cudaMemcpy(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      for (int c : {0, sms - 1}) {
        const long long* d = &h[8 * c];
        printf("  cta %3d: loop cycles %lld  wait TMA %lld  wait accumulator %lld  mmas %lld  -> %.1f cyc/mma, wall %.1f us => %.0f MHz\n", c, d[0], d[1], d[4],
               d[2], d[2] ? (double)d[0] / d[2] : 0.0, d[3] / 1e3, d[3] ? d[0] / (d[3] / 1e3) : 0.0);
      }
    }
  }
  CK(cudaGetLastError());
  const double us = ms * 1e3 / reps;
  printf("KBENCH %s N=%d %dx%dx%d Cin=%d+%d Cout=%d prec=%d ks=%d : %.1f us/launch  %.1f TFLOP/s (algorithmic)\n", op.c_str(), N, D, H, W, Cin, cin2,
         Cout, prec, ks, us, flops / (us * 1e-6) / 1e12);
  return 0;
}
