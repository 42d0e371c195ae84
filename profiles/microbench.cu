// microbench.cu -- measured per-GPU ceilings for the roofs the NDT kernels actually sit under (the driver's
// MEASURED_PEAKS.json has HBM copy bandwidth and bf16 tensor throughput only):
//   fp64 FMA/s, fp32 FMA/s (scalar and packed f32x2), integer IMAD/s = warp-instruction issue rate,
//   shared-memory read bandwidth, L2 read bandwidth (working set 32 MB), L1-resident gather rate (random 8-byte loads
//   from a 1.3 MB table), local HBM read bandwidth (2 GB stream).
// Also a numerics probe: are mul.rn.f32x2 -> add.rn.f32x2 chains kept unfused (bit-identical to __fmul_rn/__fadd_rn)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o profiles/microbench profiles/microbench.cu
// Run on the GPU box; prints one JSON object (saved as profiles/peaks_extra.json).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void mul2(float &rx, float &ry, float ax, float ay, float bx, float by) {
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
      : "=f"(rx), "=f"(ry) : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}
__device__ __forceinline__ void add2(float &rx, float &ry, float ax, float ay, float bx, float by) {
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
      : "=f"(rx), "=f"(ry) : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}
__device__ __forceinline__ void fma2(float &rx, float &ry, float ax, float ay, float bx, float by, float cx, float cy) {
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
      : "=f"(rx), "=f"(ry) : "f"(ax), "f"(ay), "f"(bx), "f"(by), "f"(cx), "f"(cy));
}

constexpr int ITER = 4096;

__global__ void k_dfma(double *out, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < ITER; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void k_ffma(float *out, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < ITER; ++i) {
    x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
    x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void k_ffma2(float *out, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < ITER; ++i) {
    fma2(x0, x1, x0, x1, a, a, b, b); fma2(x2, x3, x2, x3, a, a, b, b);
    fma2(x4, x5, x4, x5, a, a, b, b); fma2(x6, x7, x6, x7, a, a, b, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void k_imad(int *out, int a, int b) {
  int x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < ITER; ++i) {
    x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
    x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
// mixed issue: alternating fma-pipe (IMAD) and alu-pipe (LOP3 / IADD3) instructions
__global__ void k_mixed(int *out, int a, int b) {
  int x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < ITER; ++i) {
    x0 = x0 * a + b; x1 = (x1 ^ a) + b; x2 = x2 * a + b; x3 = (x3 ^ a) + b;
    x4 = x4 * a + b; x5 = (x5 ^ a) + b; x6 = x6 * a + b; x7 = (x7 ^ a) + b;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void k_smem(float *out, int iters) {
  extern __shared__ float4 s4[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s4[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  float4 acc = make_float4(0, 0, 0, 0);
  int idx = threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) { const float4 v = s4[(idx + u * 256) & 4095]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    idx = (idx + 37) & 4095;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}
__global__ void k_read(const float4 *__restrict__ p, size_t n4, int reps, float *out) {
  float4 acc = make_float4(0, 0, 0, 0);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (i + u * stride < n4) ? __ldcg(p + i + u * stride) : make_float4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}
// random 8-byte gathers from a table (the centroid probe of the NDT kernels): 9 loads issued back to back per step
__global__ void k_gather(const float2 *__restrict__ tab, int n_tab_mask, int iters, float *out) {
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    float2 v[9];
    s = s * 1664525u + 1013904223u;
    const int base = (s >> 8) & n_tab_mask;
#pragma unroll
    for (int u = 0; u < 9; ++u) v[u] = __ldg(tab + ((base + (u / 3) * 406 + (u % 3)) & n_tab_mask));
#pragma unroll
    for (int u = 0; u < 9; ++u) acc += v[u].x * v[u].y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// numerics probe: packed chains vs scalar _rn intrinsics on values chosen so that a fused multiply-add differs
__global__ void k_probe(const float *in, int n, int *mismatch, int *fused_would_differ) {
  int bad = 0, differ = 0, bad_sp = 0, bad_d = 0, bad_cell = 0;
  for (int i = threadIdx.x; i + 5 < n; i += blockDim.x) {
    const float c = in[i], s = in[i + 1], x = in[i + 2], y = in[i + 3], tx = in[i + 4], ty = in[i + 5];
    // scalar reference: x' = (c x + (-s) y) + tx ; y' = (s x + c y) + ty
    const float rx = __fadd_rn(__fadd_rn(__fmul_rn(c, x), __fmul_rn(-s, y)), tx);
    const float ry = __fadd_rn(__fadd_rn(__fmul_rn(s, x), __fmul_rn(c, y)), ty);
    float ax, ay, bx, by, qx, qy;
    mul2(ax, ay, c, s, x, x);            // (c x, s x)
    mul2(bx, by, -s, c, y, y);           // (-s y, c y)
    add2(qx, qy, ax, ay, bx, by);
    add2(qx, qy, qx, qy, tx, ty);
    if (__float_as_int(qx) != __float_as_int(rx) || __float_as_int(qy) != __float_as_int(ry)) ++bad;
    const float fx = __fadd_rn(fmaf(c, x, __fmul_rn(-s, y)), tx);
    if (__float_as_int(fx) != __float_as_int(rx)) ++differ;
    // variant: scalar .rn multiplies, packed adds
    float ux, uy, vx, vy;
    add2(ux, uy, __fmul_rn(c, x), __fmul_rn(s, x), __fmul_rn(-s, y), __fmul_rn(c, y));
    add2(vx, vy, ux, uy, tx, ty);
    if (__float_as_int(vx) != __float_as_int(rx) || __float_as_int(vy) != __float_as_int(ry)) ++bad_sp;
    // squared distance: packed subtract, scalar squares, scalar add  vs  all scalar
    const float dx = __fsub_rn(rx, tx), dy = __fsub_rn(ry, ty);
    const float dref = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    float ex, ey;
    add2(ex, ey, vx, vy, -tx, -ty);
    const float dd = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
    if (__float_as_int(dd) != __float_as_int(dref)) ++bad_d;
    // cell index: float floor / subtract / truncate  vs  cvt.rmi + integer subtract (|index| < 2^22)
    const float inv = 2.0f; const int min_b = -37;
    const int c_ref = (int)__fsub_rn(floorf(__fmul_rn(rx, inv)), (float)min_b);
    const int c_fast = __float2int_rd(__fmul_rn(rx, inv)) - min_b;
    if (c_ref != c_fast) ++bad_cell;
  }
  atomicAdd(mismatch, bad);
  atomicAdd(fused_would_differ, differ);
  atomicAdd(mismatch + 2, bad_sp);
  atomicAdd(mismatch + 3, bad_d);
  atomicAdd(mismatch + 4, bad_cell);
}

template <class F> static float time_ms(F launch, int reps = 5) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); best = std::min(best, ms);
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  void *scratch; CK(cudaMalloc(&scratch, (size_t)sms * 8 * 1024 * 8));
  const int blocks = sms * 8, threads = 256;
  const double nthr = (double)blocks * threads;
  float ms;
  ms = time_ms([&] { k_dfma<<<blocks, threads>>>((double *)scratch, 1.0000001, 1e-9); });
  const double dfma = nthr * ITER * 8 / (ms * 1e-3);
  ms = time_ms([&] { k_ffma<<<blocks, threads>>>((float *)scratch, 1.0000001f, 1e-9f); });
  const double ffma = nthr * ITER * 8 / (ms * 1e-3);
  ms = time_ms([&] { k_ffma2<<<blocks, threads>>>((float *)scratch, 1.0000001f, 1e-9f); });
  const double ffma2 = nthr * ITER * 8 / (ms * 1e-3);
  ms = time_ms([&] { k_imad<<<blocks, threads>>>((int *)scratch, 3, 7); });
  const double imad = nthr * ITER * 8 / (ms * 1e-3);
  ms = time_ms([&] { k_mixed<<<blocks, threads>>>((int *)scratch, 3, 7); });
  const double mixed = nthr * ITER * 12 / (ms * 1e-3);     // 4 IMAD + 4 LOP3 + 4 IADD per iteration
  CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const int sm_iters = 2048;
  ms = time_ms([&] { k_smem<<<sms * 2, 256, 65536>>>((float *)scratch, sm_iters); });
  const double smem_gbs = (double)sms * 2 * 256 * sm_iters * 8 * 16 / (ms * 1e-3) / 1e9;
  // L2-resident stream (32 MB, read 16 times) and HBM stream (2 GB)
  float4 *buf; const size_t big = (size_t)2 << 30; CK(cudaMalloc(&buf, big)); CK(cudaMemset(buf, 1, big));
  const size_t l2n = ((size_t)32 << 20) / 16;
  ms = time_ms([&] { k_read<<<sms * 8, 256>>>(buf, l2n, 16, (float *)scratch); });
  const double l2_gbs = (double)l2n * 16 * 16 / (ms * 1e-3) / 1e9;
  ms = time_ms([&] { k_read<<<sms * 8, 256>>>(buf, big / 16, 1, (float *)scratch); });
  const double hbm_gbs = (double)big / (ms * 1e-3) / 1e9;
  // L1/L2 gather: 1.3 MB table (the C4 probe table: 406 x 406 float2), 16 warps per SM like k_align_warp and 64
  const int g_iters = 4096;
  ms = time_ms([&] { k_gather<<<sms * 2, 256>>>((const float2 *)buf, (1 << 17) - 1, g_iters, (float *)scratch); });
  const double gather16 = (double)sms * 2 * 256 * g_iters * 9 / (ms * 1e-3);
  ms = time_ms([&] { k_gather<<<sms * 8, 256>>>((const float2 *)buf, (1 << 17) - 1, g_iters, (float *)scratch); });
  const double gather64 = (double)sms * 8 * 256 * g_iters * 9 / (ms * 1e-3);
  // numerics probe
  const int np = 1 << 16;
  std::vector<float> h(np);
  srand(7);
  for (int i = 0; i < np; ++i) h[i] = (float)((rand() / (double)RAND_MAX - 0.5) * 200.0);
  float *d_in; int *d_cnt; CK(cudaMalloc(&d_in, np * 4)); CK(cudaMalloc(&d_cnt, 32)); CK(cudaMemset(d_cnt, 0, 32));
  CK(cudaMemcpy(d_in, h.data(), np * 4, cudaMemcpyHostToDevice));
  k_probe<<<1, 1024>>>(d_in, np, d_cnt, d_cnt + 1);
  int cnt[8]; CK(cudaMemcpy(cnt, d_cnt, 32, cudaMemcpyDeviceToHost));
  const double issue_peak = (double)sms * 4 * clk_khz * 1e3;      // 4 schedulers x 1 warp instruction / clock
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz_max\": %.0f,\n"
         " \"fp64_fma_per_s\": %.4e, \"fp64_tflops\": %.2f,\n"
         " \"fp32_fma_per_s\": %.4e, \"fp32_tflops\": %.2f, \"fp32x2_fma_per_s\": %.4e, \"fp32x2_tflops\": %.2f,\n"
         " \"imad_thread_inst_per_s\": %.4e, \"imad_warp_inst_per_s\": %.4e, \"mixed_warp_inst_per_s\": %.4e,\n"
         " \"issue_peak_warp_inst_per_s_nominal\": %.4e,\n"
         " \"smem_read_gbs\": %.1f, \"l2_read_gbs_32mb\": %.1f, \"hbm_read_gbs_2gb\": %.1f,\n"
         " \"gather_8B_loads_per_s_16warps_sm\": %.4e, \"gather_8B_loads_per_s_64warps_sm\": %.4e,\n"
         " \"f32x2_probe\": {\"chains\": %d, \"packed_mul_add_vs_scalar_rn_mismatches\": %d, \"fused_would_differ\": %d,\n"
         "  \"scalar_mul_packed_add_mismatches\": %d, \"packed_sub_dist_mismatches\": %d, \"cvt_rmi_cell_mismatches\": %d}}\n",
         prop.name, sms, clk_khz / 1e3, dfma, 2 * dfma / 1e12, ffma, 2 * ffma / 1e12, ffma2, 2 * ffma2 / 1e12, imad, imad / 32, mixed / 32,
         issue_peak, smem_gbs, l2_gbs, hbm_gbs, gather16, gather64, np - 5, cnt[0], cnt[1], cnt[2], cnt[3], cnt[4]);
  return 0;
}
