// microbench.cu -- hardware facts behind the design decisions of the quadrature kernels:
//   1. FP64 FMA (DFMA) vs FP64 tensor (mma.sync.m8n8k4.f64 = DMMA) throughput   -> north-star "DMMA only if it wins"
//   2. red.global.add.f64 throughput for the scatter's address patterns           -> atomics vs shared-memory accumulation
//   3. cp.async.bulk + mbarrier round trip (the staging primitive of quad_sf3)    -> PTX sanity before building on it
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench.bin tools/microbench.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
  for (int i = 0; i < iters; i++)
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int k = 0; k < 8; k++) x[k] = fma(x[k], a, b);
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += x[k];
  if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int k = 0; k < NACC; k++) { c[k][0] = threadIdx.x + k; c[k][1] = k; }
  for (int i = 0; i < iters; i++)
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int k = 0; k < NACC; k++) dmma884(c[k][0], c[k][1], a, b);
  double s = 0;
#pragma unroll
  for (int k = 0; k < NACC; k++) s += c[k][0] + c[k][1];
  if (s == 12345.678) out[0] = s;
}

// DMMA numerics check: D = A(8x4) * B(4x8) with the documented fragment layout
__global__ void dmma_check_kernel(const double* A, const double* B, double* D) {
  const int l = threadIdx.x;
  double c0 = 0, c1 = 0;
  dmma884(c0, c1, A[(l / 4) * 4 + (l % 4)], B[(l % 4) * 8 + (l / 4)]);
  D[(l / 4) * 8 + 2 * (l % 4)] = c0;
  D[(l / 4) * 8 + 2 * (l % 4) + 1] = c1;
}

// pattern 0: a warp adds to 32 contiguous doubles at a pseudo-random 256B-aligned place
// pattern 1: 8 groups of 4 contiguous doubles (32B sectors) at 8 random places  (the v2 kernel's scatter)
// pattern 2: 32 random places
// pattern 3: like 0 but plain stores (reference for the LSU path without the L2 atomic unit)
__global__ void __launch_bounds__(256) red_kernel(double* buf, size_t n, int iters, int pattern) {
  const int lane = threadIdx.x & 31;
  uint64_t s = (blockIdx.x * 8ull + (threadIdx.x >> 5)) * 0x9E3779B97F4A7C15ull + 12345;
  for (int i = 0; i < iters; i++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    size_t idx;
    if (pattern == 0 || pattern == 3) idx = ((s >> 20) % (n / 32)) * 32 + lane;
    else if (pattern == 1) { uint64_t t = s + (lane >> 2) * 0x632BE59BD9B4E019ull; t ^= t >> 29; t *= 0xBF58476D1CE4E5B9ull; idx = ((t >> 20) % (n / 4)) * 4 + (lane & 3); }
    else { uint64_t t = s + lane * 0x632BE59BD9B4E019ull; t ^= t >> 29; t *= 0xBF58476D1CE4E5B9ull; idx = (t >> 20) % n; }
    if (pattern == 3) buf[idx] = 1.0;
    else atomicAdd(buf + idx, 1.0);
  }
}

// same question with cheap addressing (the u64 modulo above costs more issue slots than the reduction itself): a warp adds to
// `width` contiguous doubles (lanes >= width idle) at pseudo-random 256B-aligned places of a power-of-two buffer; mode 0 = red,
// 1 = plain store, 2 = red of 4 consecutive 256B lines per address computation (a row's 112-entry flush)
__global__ void __launch_bounds__(256) red2_kernel(double* buf, uint32_t mask32, int iters, int width, int mode) {
  const int lane = threadIdx.x & 31;
  uint32_t s = (blockIdx.x * 8u + (threadIdx.x >> 5)) * 0x9E3779B9u + 12345u;
  const bool on = lane < width;
  for (int i = 0; i < iters; i++) {
    s = s * 1664525u + 1013904223u;
    const uint32_t place = ((s >> 7) ^ (s << 9)) & mask32;
    double* p = buf + (size_t)place * 32 + lane;
    if (mode == 2) {
      if (on) { atomicAdd(p, 1.0); atomicAdd(p + 32, 1.0); atomicAdd(p + 64, 1.0); atomicAdd(p + 96, 1.0); }
    } else if (on) {
      if (mode == 1) *p = 1.0; else atomicAdd(p, 1.0);
    }
  }
}

// shared-memory accumulation step of the planned kernel: S[idx] += v, conflict-free, vs ATOMS
__global__ void __launch_bounds__(256) smem_acc_kernel(double* out, int iters) {
  __shared__ double S[5888];
  for (int t = threadIdx.x; t < 5888; t += 256) S[t] = 0;
  __syncthreads();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) { const int idx = (k * 256 + threadIdx.x + i) % 5888; S[idx] += 1.0; }
    __syncthreads();
  }
  if (S[threadIdx.x] == -1.0) out[0] = 1;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one thread arms an mbarrier with the byte count and issues a 1-D bulk copy; everyone waits on the barrier phase
__global__ void bulk_kernel(const double* src, double* dst, int ndoubles, int rounds) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* buf = reinterpret_cast<double*>(smem_raw);                    // two stages
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + 2 * ndoubles * 8);
  const uint32_t bytes = ndoubles * 8;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase[2] = {0, 0};
  double acc = 0;
  for (int r = 0; r < rounds; r++) {
    const int st = r & 1;
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[st])), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(buf + st * ndoubles)), "l"(src + (size_t)(blockIdx.x * rounds + r) * ndoubles), "r"(bytes), "r"(smem_u32(&bar[st])) : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar[st])), "r"(phase[st]) : "memory");
    phase[st] ^= 1;
    for (int t = threadIdx.x; t < ndoubles; t += blockDim.x) acc += buf[st * ndoubles + t];
    __syncthreads();   // everyone done with the stage before it is refilled two rounds later
  }
  atomicAdd(dst + blockIdx.x, acc);
}

static float time_ms(cudaEvent_t e0, cudaEvent_t e1) { float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms; }

int main() {
  int sms = 0, clk = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  printf("{\"sms\": %d, \"clock_khz\": %d", sms, clk);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double* d; CK(cudaMalloc(&d, 1 << 20));
  // ---- 1. DFMA vs DMMA ----
  {
    const int iters = 2048, blocks = sms * 4;
    double best = 0;
    for (int it = 0; it < 6; it++) {
      CK(cudaEventRecord(e0)); dfma_kernel<<<blocks, 256>>>(d, iters, 0.999999, 1e-9); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      double tf = 2.0 * 64 * iters * 256.0 * blocks / (time_ms(e0, e1) * 1e-3) / 1e12;
      if (it >= 1 && tf > best) best = tf;
    }
    printf(", \"dfma_tflops\": %.2f", best);
    for (int cfg = 0; cfg < 6; cfg++) {
      const int nacc = (cfg % 3 == 0) ? 2 : (cfg % 3 == 1 ? 4 : 8), bl = sms * (cfg < 3 ? 2 : 4);
      best = 0;
      for (int it = 0; it < 6; it++) {
        CK(cudaEventRecord(e0));
        if (nacc == 2) dmma_kernel<2><<<bl, 256>>>(d, iters, 0.999999, 1e-9);
        else if (nacc == 4) dmma_kernel<4><<<bl, 256>>>(d, iters, 0.999999, 1e-9);
        else dmma_kernel<8><<<bl, 256>>>(d, iters, 0.999999, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        // one m8n8k4 = 8*8*4 FMA = 512 flop per warp instruction
        double tf = 512.0 * 8 * nacc * iters * 8.0 * bl / (time_ms(e0, e1) * 1e-3) / 1e12;
        if (it >= 1 && tf > best) best = tf;
      }
      printf(", \"dmma_tflops_acc%d_cta%d\": %.2f", nacc, cfg < 3 ? 2 : 4, best);
    }
    // numerics
    std::vector<double> A(32), B(32), D(64), R(64, 0.0);
    for (int i = 0; i < 32; i++) { A[i] = 0.5 + i * 0.25; B[i] = 1.0 - i * 0.125; }
    for (int m = 0; m < 8; m++) for (int n = 0; n < 8; n++) for (int k = 0; k < 4; k++) R[m * 8 + n] += A[m * 4 + k] * B[k * 8 + n];
    double *dA, *dB, *dD; CK(cudaMalloc(&dA, 256)); CK(cudaMalloc(&dB, 256)); CK(cudaMalloc(&dD, 512));
    CK(cudaMemcpy(dA, A.data(), 256, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), 256, cudaMemcpyHostToDevice));
    dmma_check_kernel<<<1, 32>>>(dA, dB, dD);
    CK(cudaMemcpy(D.data(), dD, 512, cudaMemcpyDeviceToHost));
    double err = 0; for (int i = 0; i < 64; i++) err = fmax(err, fabs(D[i] - R[i]));
    printf(", \"dmma_layout_maxerr\": %.3g", err);
  }
  // ---- 2. red.f64 ----
  {
    const size_t n = (size_t)768 << 20;   // 6 GiB of doubles: far beyond L2, like the cfg-2 value array
    double* buf; CK(cudaMalloc(&buf, n * 8)); CK(cudaMemset(buf, 0, n * 8));
    const int iters = 2000, blocks = sms * 8;
    for (int pat = 0; pat < 4; pat++) {
      float best = 1e30f;
      for (int it = 0; it < 3; it++) {
        CK(cudaEventRecord(e0)); red_kernel<<<blocks, 256>>>(buf, n, iters, pat); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        best = fminf(best, time_ms(e0, e1));
      }
      const double ops = (double)blocks * 256 * iters;
      printf(", \"red_pattern%d_Gops\": %.2f", pat, ops / (best * 1e-3) / 1e9);
    }
    // same with an L2-resident target (64 MiB)
    for (int pat = 0; pat < 3; pat++) {
      float best = 1e30f;
      for (int it = 0; it < 3; it++) {
        CK(cudaEventRecord(e0)); red_kernel<<<blocks, 256>>>(buf, (size_t)8 << 20, iters, pat); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        best = fminf(best, time_ms(e0, e1));
      }
      printf(", \"red_l2_pattern%d_Gops\": %.2f", pat, (double)blocks * 256 * iters / (best * 1e-3) / 1e9);
    }
    // cheap-address versions: 4 GiB target (mask over 2^24 256-byte lines) and an L2-resident 64 MiB target
    for (int l2 = 0; l2 < 2; l2++)
      for (int cfg = 0; cfg < 5; cfg++) {
        const int width = (cfg == 3) ? 28 : 32, mode = (cfg == 1) ? 1 : (cfg == 2 ? 2 : 0);
        const int bl = (cfg == 4) ? sms : blocks;                   // cfg 4: one CTA of 8 warps per SM (the occupancy of quad_sf3)
        const uint32_t mask = l2 ? ((1u << 18) - 1) : ((1u << 24) - 1);
        float best = 1e30f;
        for (int it = 0; it < 3; it++) {
          CK(cudaEventRecord(e0)); red2_kernel<<<bl, 256>>>(buf, mask, iters, width, mode); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
          best = fminf(best, time_ms(e0, e1));
        }
        const double ops = (double)bl * 8 * width * iters * (mode == 2 ? 4 : 1);
        static const char* nm[5] = {"red32", "store32", "red4x32", "red28", "red32_1cta"};
        printf(", \"%s_%s_Gops\": %.2f", nm[cfg], l2 ? "l2" : "dram", ops / (best * 1e-3) / 1e9);
      }
    CK(cudaFree(buf));
  }
  // ---- smem accumulate ----
  {
    const int iters = 4000, blocks = sms * 3;
    float best = 1e30f;
    for (int it = 0; it < 3; it++) {
      CK(cudaEventRecord(e0)); smem_acc_kernel<<<blocks, 256>>>(d, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      best = fminf(best, time_ms(e0, e1));
    }
    printf(", \"smem_acc_Gops\": %.2f", (double)blocks * 256 * 16 * iters / (best * 1e-3) / 1e9);
  }
  // ---- 3. cp.async.bulk + mbarrier ----
  {
    const int nd = 576, rounds = 64, blocks = sms * 2;     // 4608-byte slices = one PP table of the p=3 kernel
    std::vector<double> h((size_t)blocks * rounds * nd);
    double expect = 0;
    for (size_t i = 0; i < h.size(); i++) { h[i] = (double)(i % 1000) * 1e-3; }
    for (size_t i = 0; i < (size_t)rounds * nd; i++) expect += h[i];
    double *src, *dst; CK(cudaMalloc(&src, h.size() * 8)); CK(cudaMalloc(&dst, blocks * 8)); CK(cudaMemset(dst, 0, blocks * 8));
    CK(cudaMemcpy(src, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    const int smem = 2 * nd * 8 + 64;
    CK(cudaEventRecord(e0)); bulk_kernel<<<blocks, 256, smem>>>(src, dst, nd, rounds); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    double got; CK(cudaMemcpy(&got, dst, 8, cudaMemcpyDeviceToHost));
    printf(", \"bulk_copy_relerr\": %.3g, \"bulk_copy_ms\": %.3f", fabs(got - expect) / expect, time_ms(e0, e1));
  }
  printf("}\n");
  return 0;
}
