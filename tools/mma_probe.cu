// mma_probe.cu -- developer microbenchmark (not part of the product): what a SMALL tcgen05.mma costs on
// B200, as seen by the issuing thread and by the tensor pipe, for the operand layouts fused_tc.cu uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/mma_probe tools/mma_probe.cu
// Every line: layout, N, CTAs/SM, #accumulators used round-robin, MMAs per commit (group), then cycles per MMA
//   issue  = clock64 around the issue loop only (thread 0)
//   total  = issue + commit + wait of the LAST commit (all groups committed, one wait at the end)
//   serial = every group is committed AND awaited before the next is issued (the dependent-stage pattern)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t it = 0; !done; ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(100000u) : "memory");
        if (it > (1u << 20)) __trap();
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory"); }
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout)
{
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
           ((uint64_t)layout << 61);
}
__device__ __forceinline__ uint32_t instr_desc(int M, int N, int a_mn, int b_mn)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Probe {
    int layout;   // 0 K-major no swizzle (slab layout), 1 MN-major no swizzle (the dW use), 2 K-major 128B swizzle, 3 A from TMEM + B K-major no swizzle
    int M, N, n_mma, n_acc, group, serial, tmem_cols, smem_bytes, b_off, always_acc, issuers;
};

__global__ void __launch_bounds__(128) probe_kernel(Probe p, long long *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < p.smem_bytes / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(smem_u32(&slot), p.tmem_cols);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if ((tid & 31) == 0 && warp < p.issuers) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + p.b_off);
        uint64_t da, db;
        uint32_t id, inc_a, inc_b;
        const int SLAB = 2048;
        if (p.layout == 0 || p.layout == 3) {        // rows at 16 B inside an 8-feature slab
            da = desc(a0, SLAB, 128, 0); db = desc(b0, p.N * 16, 128, 0);
            id = instr_desc(p.M, p.N, 0, 0); inc_a = (2 * SLAB) >> 4; inc_b = (2 * p.N * 16) >> 4;
        } else if (p.layout == 1) {                  // the same bytes read MN-major (rows = K)
            da = desc(a0, 128, SLAB, 0); db = desc(b0, 128, SLAB, 0);
            id = instr_desc(p.M, p.N, 1, 1); inc_a = 256 >> 4; inc_b = 256 >> 4;
        } else {                                     // 128-byte swizzle, K-major, 64-element rows
            da = desc(a0, 16, 1024, 2); db = desc(b0, 16, 1024, 2);
            id = instr_desc(p.M, p.N, 0, 0); inc_a = 32 >> 4; inc_b = 32 >> 4;
        }
        const uint32_t acc_stride = (uint32_t)((p.N + 31) / 32 * 32);
        const uint32_t a_tmem = tmem + (uint32_t)p.tmem_cols - 32;   // 8 columns of "A" per K step (4 steps)
        uint32_t phase = 0;
        long long best_issue = 1ll << 60, best_total = 1ll << 60;
        for (int rep = 0; rep < 5; ++rep) {
            const long long t0 = clock64();
            long long t_issue = 0;
            for (int i = 0; i < p.n_mma; i += p.group) {
                const long long ti = p.serial ? clock64() : 0;
                const uint32_t d = tmem + (uint32_t)(((i / p.group) % p.n_acc) + warp * p.n_acc) * acc_stride;
                for (int k = 0; k < p.group; ++k) {
                    const uint64_t a = da + (uint64_t)(inc_a * (uint32_t)(k & 3)), b = db + (uint64_t)(inc_b * (uint32_t)(k & 3));
                    if (p.layout == 3) umma_ts(d, a_tmem + (uint32_t)(k & 3) * 8, b, id, (k > 0) | p.always_acc);
                    else umma_ss(d, a, b, id, (k > 0) | p.always_acc);
                }
                if (p.serial) t_issue += clock64() - ti;
                if (p.serial || i + p.group >= p.n_mma) {
                    umma_commit(smem_u32(&bar[warp]));
                    mbar_wait(smem_u32(&bar[warp]), phase);
                    phase ^= 1;
                    tc_fence_after();
                }
            }
            const long long t1 = clock64();
            if (!p.serial) t_issue = t1 - t0;
            if (rep > 0) { best_issue = t_issue < best_issue ? t_issue : best_issue; best_total = (t1 - t0) < best_total ? (t1 - t0) : best_total; }
        }
        if (warp == 0) { out[blockIdx.x * 2] = best_issue; out[blockIdx.x * 2 + 1] = best_total; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, p.tmem_cols);
}

int main()
{
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    long long *d_out;
    cudaMalloc(&d_out, sizeof(long long) * 2 * sm * 8);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    const char *lname[4] = {"K-major/none", "MN-major/none", "K-major/sw128", "A=TMEM,B=K/none"};
    printf("%-16s %4s %4s %5s %5s %6s | %9s %9s   (cycles per MMA, median over CTAs)\n", "layout", "M", "N", "cta/sm", "n_acc", "group", "issue", "total");
    auto run = [&](int layout, int M, int N, int per_sm, int n_acc, int group, int serial, int always_acc = 0, int issuers = 1) {
        Probe p{layout, M, N, 96, n_acc, group, serial, 512 / per_sm >= 512 ? 512 : (512 / per_sm), per_sm == 1 ? 96 * 1024 : 48 * 1024, per_sm == 1 ? 48 * 1024 : 16 * 1024, always_acc, issuers};
        if (p.tmem_cols > 256 && per_sm > 1) p.tmem_cols = 256;
        const int need = ((N + 31) / 32 * 32) * n_acc * issuers + 32;
        if (need > p.tmem_cols) return;
        const int b_bytes = layout == 1 ? (N / 8) * 2048 : (layout == 2 ? N * 128 : N * 32 * 4);
        if (p.b_off + b_bytes > p.smem_bytes) return;
        p.n_mma = 96 / group * group;
        const int grid = sm * per_sm;
        probe_kernel<<<grid, 128, p.smem_bytes>>>(p, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-16s N=%d: %s\n", lname[layout], N, cudaGetErrorString(e)); exit(1); }
        std::vector<long long> h(2 * grid);
        cudaMemcpy(h.data(), d_out, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
        std::vector<double> is, to;
        for (int b = 0; b < grid; ++b) { is.push_back((double)h[2 * b] / p.n_mma); to.push_back((double)h[2 * b + 1] / p.n_mma); }
        std::sort(is.begin(), is.end()); std::sort(to.begin(), to.end());
        printf("%-16s %4d %4d %5d %5d %6d | %9.1f %9.1f   %s (x cta/sm = %.1f per SM)\n", lname[layout], M, N, per_sm, n_acc, group, is[grid / 2], to[grid / 2],
               serial ? "serial" : "piped ", to[grid / 2] / per_sm); if (always_acc || issuers > 1) printf("      ^ always_acc=%d issuers=%d\n", always_acc, issuers);
    };
    // A. what does starting a chain cost?  serial stages with the first MMA overwriting vs always accumulating
    for (int layout : {0, 3})
        for (int group : {1, 2, 3}) { run(layout, 128, 32, 1, 1, group, 1, 0); run(layout, 128, 32, 1, 1, group, 1, 1); }
    // B. alternating accumulators without waits: overwrite vs accumulate
    for (int group : {1, 2}) { run(0, 128, 32, 1, 4, group, 0, 0); run(0, 128, 32, 1, 4, group, 0, 1); }
    // C. several issuing warps in ONE CTA, each its own accumulator and barrier
    for (int issuers : {1, 2, 4}) { run(0, 128, 32, 1, 1, 96, 0, 0, issuers); run(0, 128, 32, 1, 1, 2, 1, 0, issuers); run(1, 128, 80, 1, 1, 8, 1, 0, issuers); }
    for (int issuers : {2, 4}) run(0, 128, 32, 2, 1, 2, 1, 0, issuers);
    return 0;
}
