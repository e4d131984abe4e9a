// NVLink store-throughput microbenchmark (2 GPUs, no torch): how fast can SM-issued stores leave the
// GPU towards a peer, as a function of the store flavour? This is the ceiling of phase A / B of the
// fused allreduce kernel (profiles/r2/p2p_store_bench.md).
//   st8 / st16 / st32 : st.global.v2/v4/v8 from registers, a warp writes one contiguous run per instruction
//   tma<N>            : the warp stages N bytes in shared memory, one lane issues cp.async.bulk (S2G)
// Every kernel ends with a system fence, so the timed region includes the drain of the stores.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o bench/micro/p2p_store_bench bench/micro/p2p_store_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      exit(1);                                                                      \
    }                                                                               \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// BYTES per lane per store instruction: 8, 16, 32
template <int BYTES>
__global__ void __launch_bounds__(256, 2) store_kernel(uint8_t* dst, size_t total, uint32_t seed) {
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (size_t)gridDim.x * 8;
  const uint32_t lane = threadIdx.x & 31u;
  constexpr size_t kRun = (size_t)BYTES * 32;
  const uint32_t v = seed + threadIdx.x;
  for (size_t off = warp * kRun; off + kRun <= total; off += nwarps * kRun) {
    uint8_t* p = dst + off + (size_t)lane * BYTES;
    if constexpr (BYTES == 4) {
      asm volatile("st.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    } else if constexpr (BYTES == 8) {
      asm volatile("st.global.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(v), "r"(v + 1) : "memory");
    } else if constexpr (BYTES == 16) {
      asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v), "r"(v + 1), "r"(v + 2), "r"(v + 3) : "memory");
    } else {
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v), "r"(v + 1), "r"(v + 2), "r"(v + 3),
                   "r"(v + 4), "r"(v + 5), "r"(v + 6), "r"(v + 7)
                   : "memory");
    }
  }
  __threadfence_system();
}

// each warp owns a CHUNK-byte staging buffer (double buffered); lanes fill it with st.shared.v4, then
// one lane issues the bulk store
template <int CHUNK>
__global__ void __launch_bounds__(256, 2) tma_kernel(uint8_t* dst, size_t total, uint32_t seed) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  uint8_t* stage = smem + (size_t)w * 2 * CHUNK;
  const size_t warp = (size_t)blockIdx.x * 8 + w, nwarps = (size_t)gridDim.x * 8;
  uint32_t buf = 0;
  for (size_t off = warp * CHUNK; off + CHUNK <= total; off += nwarps * CHUNK, buf ^= 1u) {
    uint8_t* s = stage + buf * CHUNK;
    // the bulk store issued two iterations ago from this buffer must have finished READING it
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    for (uint32_t b = lane * 16; b < CHUNK; b += 32 * 16) {
      const uint32_t v = seed + b;
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(s + b)), "r"(v), "r"(v + 1), "r"(v + 2), "r"(v + 3)
                   : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(smem_u32(s)),
                   "r"((uint32_t)CHUNK)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncwarp();
  __threadfence_system();
}

// st8 plus one 8-byte "meta" store per 256-byte run into a separate region (the traffic shape of
// phase A: a packed slice + its {unit, min} record)
__global__ void __launch_bounds__(256, 2) store_meta_kernel(uint8_t* dst, uint8_t* meta, size_t total, uint32_t seed) {
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (size_t)gridDim.x * 8;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t v = seed + threadIdx.x;
  for (size_t off = warp * 256; off + 256 <= total; off += nwarps * 256) {
    asm volatile("st.global.v2.b32 [%0], {%1,%2};" ::"l"(dst + off + (size_t)lane * 8), "r"(v), "r"(v + 1) : "memory");
    if (lane == 0) asm volatile("st.global.v2.b32 [%0], {%1,%2};" ::"l"(meta + (off >> 8) * 8), "r"(v), "r"(v) : "memory");
  }
  __threadfence_system();
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// The release at the end of a phase, in isolation: every CTA stores `per_cta` bytes to the peer
// (8-byte stores), then publishes a flag. MODE 0: bar.sync, warp 0 fences (what the fused kernel
// does); MODE 1: every warp fences its own stores first, then bar.sync, then warp 0 fences;
// MODE 2: no fence at all (lower bound, not a valid protocol). times[cta] = {start, stores issued, released}.
template <int MODE>
__global__ void __launch_bounds__(256, 2) release_kernel(uint8_t* dst, size_t per_cta, uint32_t* flag, uint32_t seed,
                                                         unsigned long long* times) {
  const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  unsigned long long t0 = gtime();
  uint8_t* base = dst + (size_t)blockIdx.x * per_cta;
  const uint32_t v = seed + threadIdx.x;
  for (size_t off = (size_t)w * 256; off + 256 <= per_cta; off += 8 * 256)
    asm volatile("st.global.v2.b32 [%0], {%1,%2};" ::"l"(base + off + (size_t)lane * 8), "r"(v), "r"(v + 1) : "memory");
  if (MODE == 1) __threadfence_system();
  __syncthreads();
  unsigned long long t1 = gtime();
  if (w == 0) {
    if (MODE != 2) __threadfence_system();
    if (lane == 0) asm volatile("st.relaxed.sys.global.b32 [%0], %1;" ::"l"(flag + blockIdx.x), "r"(seed) : "memory");
    unsigned long long t2 = gtime();
    if (lane == 0) {
      times[blockIdx.x * 3 + 0] = t0;
      times[blockIdx.x * 3 + 1] = t1;
      times[blockIdx.x * 3 + 2] = t2;
    }
  }
}

// pull: ld.global.v4 from the peer, sum into a register, write one word locally
__global__ void __launch_bounds__(256, 2) load_kernel(const uint8_t* src, size_t total, uint32_t* sink) {
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (size_t)gridDim.x * 8;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t acc = 0;
  for (size_t off = warp * 512; off + 512 <= total; off += nwarps * 512) {
    uint32_t a, b, c, d;
    asm volatile("ld.relaxed.sys.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(src + off + lane * 16));
    acc += a ^ b ^ c ^ d;
  }
  if (acc == 0x12345678u) *sink = acc;
}

struct Result {
  const char* name;
  const char* where;
  int grid;
  double mb, us, gbs;
};

template <typename F>
double time_us(F&& launch, int iters) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < iters; ++i) launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  return (double)ms * 1000.0 / iters;
}

int main(int argc, char** argv) {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  const char* out = argc > 1 ? argv[1] : "gpurun_out/p2p_store_bench.json";
  CK(cudaSetDevice(0));
  uint8_t *local = nullptr, *peer = nullptr;
  const size_t cap = 256u << 20;
  CK(cudaMalloc(&local, cap));
  if (ndev > 1) {
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    if (can) {
      CK(cudaSetDevice(1));
      CK(cudaMalloc(&peer, cap));
      CK(cudaSetDevice(0));
      CK(cudaDeviceEnablePeerAccess(1, 0));
    }
  }
  uint32_t* sink = nullptr;
  CK(cudaMalloc(&sink, 4));
  CK(cudaFuncSetAttribute(tma_kernel<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 4096));
  std::vector<Result> res;
  const size_t sizes[] = {8u << 20, 64u << 20};
  const int grids[] = {148, 296};
  for (int where = 0; where < 2; ++where) {
    uint8_t* dst = where == 0 ? local : peer;
    if (!dst) continue;
    const char* wn = where == 0 ? "local" : "peer";
    for (size_t total : sizes) {
      for (int grid : grids) {
        const int iters = total > (16u << 20) ? 10 : 30;
        auto add = [&](const char* name, double us) {
          res.push_back({name, wn, grid, total / 1048576.0, us, total / us / 1e3});
          printf("%-8s %-5s grid %3d %6.0f MB %9.2f us %8.1f GB/s\n", name, wn, grid, total / 1048576.0, us, total / us / 1e3);
          fflush(stdout);
        };
        add("st4", time_us([&] { store_kernel<4><<<grid, 256>>>(dst, total, 1); }, iters));
        add("st8", time_us([&] { store_kernel<8><<<grid, 256>>>(dst, total, 1); }, iters));
        add("st16", time_us([&] { store_kernel<16><<<grid, 256>>>(dst, total, 1); }, iters));
        add("st32", time_us([&] { store_kernel<32><<<grid, 256>>>(dst, total, 1); }, iters));
        add("tma256", time_us([&] { tma_kernel<256><<<grid, 256, 8 * 2 * 256>>>(dst, total, 1); }, iters));
        add("tma1k", time_us([&] { tma_kernel<1024><<<grid, 256, 8 * 2 * 1024>>>(dst, total, 1); }, iters));
        add("tma4k", time_us([&] { tma_kernel<4096><<<grid, 256, 8 * 2 * 4096>>>(dst, total, 1); }, iters));
        add("st8+meta", time_us([&] { store_meta_kernel<<<grid, 256>>>(dst, dst + total, total, 1); }, iters));
        add("ld16", time_us([&] { load_kernel<<<grid, 256>>>(dst, total, sink); }, iters));
        CK(cudaGetLastError());
      }
    }
  }
  // the release in isolation (phase-A shape: 296 CTAs x 24 KB = 7.1 MB)
  if (peer) {
    const int grid = 296;
    unsigned long long* d_times = nullptr;
    CK(cudaMalloc(&d_times, grid * 3 * sizeof(unsigned long long)));
    std::vector<unsigned long long> h(grid * 3);
    const size_t per_ctas[] = {24u << 10, 96u << 10};
    for (size_t per_cta : per_ctas) {
      for (int mode = 0; mode < 3; ++mode) {
        double s_store = 0, s_rel = 0, s_all = 0;
        const int reps = 20;
        for (int it = 0; it < reps + 2; ++it) {
          uint32_t* flag = reinterpret_cast<uint32_t*>(peer + cap - (1u << 20));
          if (mode == 0) release_kernel<0><<<grid, 256>>>(peer, per_cta, flag, it, d_times);
          if (mode == 1) release_kernel<1><<<grid, 256>>>(peer, per_cta, flag, it, d_times);
          if (mode == 2) release_kernel<2><<<grid, 256>>>(peer, per_cta, flag, it, d_times);
          CK(cudaMemcpy(h.data(), d_times, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
          if (it < 2) continue;
          unsigned long long first = ~0ull, last = 0;
          double a = 0, b = 0;
          for (int c = 0; c < grid; ++c) {
            a += (double)(h[c * 3 + 1] - h[c * 3 + 0]);
            b += (double)(h[c * 3 + 2] - h[c * 3 + 1]);
            first = h[c * 3] < first ? h[c * 3] : first;
            last = h[c * 3 + 2] > last ? h[c * 3 + 2] : last;
          }
          s_store += a / grid / 1e3;
          s_rel += b / grid / 1e3;
          s_all += (double)(last - first) / 1e3;
        }
        const char* names[] = {"rel_warp0", "rel_allwarps", "rel_none"};
        printf("%-12s %3zu KB/CTA: stores issued %.2f us, release %.2f us, first start .. last release %.2f us\n", names[mode],
               per_cta >> 10, s_store / reps, s_rel / reps, s_all / reps);
        res.push_back({names[mode], "peer", grid, per_cta * grid / 1048576.0, s_rel / reps, s_all / reps});
      }
    }
  }
  // copy engine for reference
  if (peer) {
    const size_t total = 64u << 20;
    double us = time_us([&] { CK(cudaMemcpyPeerAsync(peer, 1, local, 0, total, 0)); }, 10);
    res.push_back({"memcpyPeer", "peer", 0, total / 1048576.0, us, total / us / 1e3});
    printf("memcpyPeer 64 MB %.2f us %.1f GB/s\n", us, total / us / 1e3);
  }
  FILE* f = fopen(out, "w");
  if (f) {
    fprintf(f, "{\"rows\": [\n");
    for (size_t i = 0; i < res.size(); ++i)
      fprintf(f, " {\"kernel\": \"%s\", \"dst\": \"%s\", \"grid\": %d, \"mb\": %.0f, \"us\": %.2f, \"gbs\": %.1f}%s\n", res[i].name,
              res[i].where, res[i].grid, res[i].mb, res[i].us, res[i].gbs, i + 1 < res.size() ? "," : "");
    fprintf(f, "]}\n");
    fclose(f);
  }
  return 0;
}
