// p2p_probe.cu — design probe for the slab transposes / halo exchanges: what does one B200 get out of its NVLink
// port with (a) copy-engine peer copies, (b) per-thread peer stores from a kernel in runs of 64 B .. 1 KB,
// (c) TMA bulk stores (cp.async.bulk shared -> peer global) from a few CTAs, each alone and next to an HBM-bound
// kernel on the same GPU; plus the latency of a flag rendezvous.  Two GPUs, one process.  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/p2p_probe tools/p2p_probe.cu && gpurun_out/p2p_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// local HBM-bound kernel: y = x + 1 over n doubles (16 B/elem of traffic)
__global__ void k_local(const double *__restrict__ x, double *__restrict__ y, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) y[i] = x[i] + 1.0;
}
// peer stores: read local src (coalesced), store to dst in runs of RUN doubles: element e of the source goes to
// (e / RUN) scattered block order (so consecutive runs land in different 4 KB regions like the wire format does)
template <int VEC>
__global__ void k_store(const double *__restrict__ src, double *__restrict__ dst, size_t n, int run, size_t nruns) {
  size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * VEC;
  const size_t st = (size_t)gridDim.x * blockDim.x * VEC;
  for (; i < n; i += st) {
    const size_t r = i / run, o = i - r * run;
    const size_t rr = (r * 2654435761ull) % nruns;   // permute the runs (odd multiplier, nruns power of two)
    if (VEC == 2) *reinterpret_cast<double2 *>(dst + rr * run + o) = *reinterpret_cast<const double2 *>(src + i);
    else dst[rr * run + o] = src[i];
  }
}
// peer loads (pull): the mirror image
__global__ void k_pull(const double *__restrict__ remote, double *__restrict__ dst, size_t n) {
  size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 2;
  const size_t st = (size_t)gridDim.x * blockDim.x * 2;
  for (; i < n; i += st) *reinterpret_cast<double2 *>(dst + i) = *reinterpret_cast<const double2 *>(remote + i);
}
// TMA bulk: each CTA loops over chunks: bulk load global(local) -> smem, bulk store smem -> global(peer)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int STAGES>
__global__ void k_bulk(const double *__restrict__ src, double *__restrict__ dst, size_t nchunks, int chunk_bytes) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const size_t ce = chunk_bytes / 8;
  size_t it = 0;
  for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, it++) {
    const int s = it % STAGES;
    const uint32_t par = (it / STAGES) & 1;
    if (it >= STAGES) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STAGES - 1) : "memory");   // stage s's store has read its smem
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[s])), "r"(chunk_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + (size_t)s * chunk_bytes)),
                 "l"(src + c * ce), "r"(chunk_bytes), "r"(s32(&bar[s])) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar[s])), "r"(par) : "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * ce), "r"(s32(sm + (size_t)s * chunk_bytes)), "r"(chunk_bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// flag ping-pong: rank r writes epoch to the peer's flag, waits for its own
__global__ void k_flag(volatile unsigned long long *mine, volatile unsigned long long *peer, unsigned long long epoch) {
  __threadfence_system();
  *peer = epoch;
  __threadfence_system();
  while (*mine < epoch) {}
}

struct Dev { int id; cudaStream_t s, s2; double *a, *b, *win; unsigned long long *flag; cudaEvent_t e0, e1, f0, f1; };

int main() {
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
  const size_t N = (size_t)1 << 27;   // 1 GiB of doubles per buffer
  Dev d[2];
  for (int r = 0; r < 2; r++) {
    d[r].id = r;
    CK(cudaSetDevice(r));
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, r, 1 - r));
    if (!can) { printf("no peer access %d -> %d\n", r, 1 - r); return 0; }
    CK(cudaDeviceEnablePeerAccess(1 - r, 0));
    CK(cudaStreamCreateWithFlags(&d[r].s, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&d[r].s2, cudaStreamNonBlocking));
    CK(cudaMalloc(&d[r].a, N * 8)); CK(cudaMalloc(&d[r].b, N * 8)); CK(cudaMalloc(&d[r].win, N * 8));
    CK(cudaMalloc(&d[r].flag, 256)); CK(cudaMemset(d[r].flag, 0, 256));
    CK(cudaMemset(d[r].a, 0, N * 8)); CK(cudaMemset(d[r].b, 0, N * 8)); CK(cudaMemset(d[r].win, 0, N * 8));
    for (cudaEvent_t *e : {&d[r].e0, &d[r].e1, &d[r].f0, &d[r].f1}) CK(cudaEventCreate(e));
  }
  auto sync_all = [&]() { for (int r = 0; r < 2; r++) { CK(cudaSetDevice(r)); CK(cudaDeviceSynchronize()); } };
  sync_all();
  // run `xfer(r)` on stream s of the ranks in `who` (bit mask), optionally `local(r)` on s2 concurrently; report GB/s
  auto run = [&](const char *name, unsigned who, size_t bytes_x, auto xfer, bool with_local) {
    for (int rep = 0; rep < 3; rep++) {
      sync_all();
      for (int r = 0; r < 2; r++) {
        if (!((who >> r) & 1)) continue;
        CK(cudaSetDevice(r));
        CK(cudaEventRecord(d[r].e0, d[r].s));
        xfer(r);
        CK(cudaEventRecord(d[r].e1, d[r].s));
        if (with_local) {
          CK(cudaEventRecord(d[r].f0, d[r].s2));
          for (int q = 0; q < 4; q++) k_local<<<148 * 8, 256, 0, d[r].s2>>>(d[r].a, d[r].b, N);
          CK(cudaEventRecord(d[r].f1, d[r].s2));
        }
      }
      sync_all();
      if (rep < 2) continue;
      for (int r = 0; r < 2; r++) {
        if (!((who >> r) & 1)) continue;
        float ms = 0, ml = 0;
        CK(cudaEventElapsedTime(&ms, d[r].e0, d[r].e1));
        printf("%-58s gpu%d  %8.3f ms  %7.1f GB/s", name, r, ms, bytes_x / (ms * 1e-3) / 1e9);
        if (with_local) { CK(cudaEventElapsedTime(&ml, d[r].f0, d[r].f1)); printf("   | local kernel next to it: %8.3f ms %7.1f GB/s", ml, 4.0 * N * 16 / (ml * 1e-3) / 1e9); }
        printf("\n");
      }
    }
  };
  // 0. local kernel alone
  run("local HBM kernel alone (4 x 2 GiB traffic)", 3, 4 * N * 16, [&](int r) { for (int q = 0; q < 4; q++) k_local<<<148 * 8, 256, 0, d[r].s>>>(d[r].a, d[r].b, N); }, false);
  // 1. copy engine
  for (size_t mb : {8, 64, 512}) {
    char nm[128];
    const size_t by = mb << 20;
    snprintf(nm, sizeof nm, "CE peer copy %zu MiB, one direction", mb);
    run(nm, 1, by, [&](int r) { CK(cudaMemcpyPeerAsync(d[1 - r].win, 1 - r, d[r].a, r, by, d[r].s)); }, false);
    snprintf(nm, sizeof nm, "CE peer copy %zu MiB, both directions", mb);
    run(nm, 3, by, [&](int r) { CK(cudaMemcpyPeerAsync(d[1 - r].win, 1 - r, d[r].a, r, by, d[r].s)); }, false);
  }
  run("CE peer copy 512 MiB both dirs + local HBM kernel", 3, (size_t)512 << 20, [&](int r) { CK(cudaMemcpyPeerAsync(d[1 - r].win, 1 - r, d[r].a, r, (size_t)512 << 20, d[r].s)); }, true);
  run("CE 8 x 64 MiB copies on one stream, both dirs", 3, (size_t)512 << 20, [&](int r) { for (int q = 0; q < 8; q++) CK(cudaMemcpyPeerAsync(d[1 - r].win + q * ((size_t)8 << 20), 1 - r, d[r].a + q * ((size_t)8 << 20), r, (size_t)64 << 20, d[r].s)); }, false);
  // 2. kernel peer stores
  const size_t NS = (size_t)1 << 26;   // 512 MiB
  for (int run_d : {8, 16, 32, 128}) {
    for (int grid : {148 * 8, 32}) {
      char nm[128];
      snprintf(nm, sizeof nm, "kernel peer stores 16 B/thread, runs of %4d B, %4d CTAs, both dirs", run_d * 8, grid);
      run(nm, 3, NS * 8, [&](int r) { k_store<2><<<grid, 256, 0, d[r].s>>>(d[r].a, d[1 - r].win, NS, run_d, NS / run_d); }, false);
    }
  }
  run("kernel peer stores 8 B/thread, runs of 64 B, 1184 CTAs, both dirs", 3, NS * 8, [&](int r) { k_store<1><<<148 * 8, 256, 0, d[r].s>>>(d[r].a, d[1 - r].win, NS, 8, NS / 8); }, false);
  run("kernel peer stores 16 B, runs 1 KB, 1184 CTAs + local HBM kernel", 3, NS * 8, [&](int r) { k_store<2><<<148 * 8, 256, 0, d[r].s>>>(d[r].a, d[1 - r].win, NS, 128, NS / 128); }, true);
  run("kernel peer stores 16 B, runs 1 KB, 32 CTAs + local HBM kernel", 3, NS * 8, [&](int r) { k_store<2><<<32, 256, 0, d[r].s>>>(d[r].a, d[1 - r].win, NS, 128, NS / 128); }, true);
  run("kernel peer LOADS 16 B/thread (pull), 1184 CTAs, both dirs", 3, NS * 8, [&](int r) { k_pull<<<148 * 8, 256, 0, d[r].s>>>(d[1 - r].a, d[r].win, NS); }, false);
  run("kernel peer LOADS 16 B/thread (pull), 1184 CTAs + local HBM kernel", 3, NS * 8, [&](int r) { k_pull<<<148 * 8, 256, 0, d[r].s>>>(d[1 - r].a, d[r].win, NS); }, true);
  // 3. TMA bulk stores
  for (int r = 0; r < 2; r++) { CK(cudaSetDevice(r)); CK(cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768)); }
  for (int cb : {4096, 16384, 32768}) {
    for (int grid : {8, 16, 32, 148}) {
      char nm[128];
      snprintf(nm, sizeof nm, "TMA bulk smem->peer, %5d B chunks, 4 stages, %3d CTAs, both dirs", cb, grid);
      run(nm, 3, NS * 8, [&](int r) { k_bulk<4><<<grid, 32, 4 * cb, d[r].s>>>(d[r].a, d[1 - r].win, NS * 8 / cb, cb); }, false);
    }
  }
  run("TMA bulk smem->peer, 16 KB chunks, 16 CTAs + local HBM kernel", 3, NS * 8, [&](int r) { k_bulk<4><<<16, 32, 4 * 16384, d[r].s>>>(d[r].a, d[1 - r].win, NS * 8 / 16384, 16384); }, true);
  run("TMA bulk smem->peer, 16 KB chunks, 32 CTAs + local HBM kernel", 3, NS * 8, [&](int r) { k_bulk<4><<<32, 32, 4 * 16384, d[r].s>>>(d[r].a, d[1 - r].win, NS * 8 / 16384, 16384); }, true);
  // 4. flag rendezvous latency: 100 epochs, each GPU launches one kernel per epoch
  {
    sync_all();
    for (int r = 0; r < 2; r++) {
      CK(cudaSetDevice(r));
      CK(cudaEventRecord(d[r].e0, d[r].s));
      for (unsigned long long e = 1; e <= 100; e++) k_flag<<<1, 1, 0, d[r].s>>>(d[r].flag, d[1 - r].flag, e);
      CK(cudaEventRecord(d[r].e1, d[r].s));
    }
    sync_all();
    for (int r = 0; r < 2; r++) { float ms; CK(cudaEventElapsedTime(&ms, d[r].e0, d[r].e1)); printf("flag rendezvous: 100 back-to-back barrier kernels       gpu%d  %8.3f ms = %.1f us each\n", r, ms, ms * 10); }
  }
  printf("done\n");
  return 0;
}
