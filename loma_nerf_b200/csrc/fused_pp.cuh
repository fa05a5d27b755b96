// fused_pp.cuh -- the fused tensor-core step, second design ("two rows per thread").  Included by fused_tc.cu
// inside its anonymous namespace (PTX wrappers, TcParams, SLAB / TILE / MAXL come from there).
//
// What was measured first (tools/mma_probe.cu, B200): ONE thread issues at most one tcgen05.mma per ~114
// cycles whatever M, N (<= 128), layout or A source; a dependent stage of g MMAs (issue, commit, mbarrier
// wake-up) costs ~555 + 145 g cycles of pure latency; chains issued by different warps or CTAs overlap
// fully until the pipe's shared-memory operand fetch (~47 cycles per M=128 x K=16 MMA) binds.  And ncu on
// the first design: issue slots 33 % busy, the top stall is `wait` (dependent fixed-latency ALU chains) --
// 128 threads that each own one sample row have no instruction-level parallelism, and shared memory allows
// only 3-4 such tiles per SM.  (An intermediate version that ping-ponged two tiles through a dynamic phase
// state machine hid the MMA latency but executed 50 % more instructions and was slower.)
//
// This design keeps the per-tile arithmetic (same slab layout, same MMA shapes) and changes the schedule:
//   * warps 0-3  "row" threads: thread r owns sample row r (TMEM lane r) of TWO tiles (slots) and takes both
//                through every phase together: two independent instruction streams per thread in the epilogues
//                and the compositing scans, one barrier / one MMA-stage wait per PAIR of tiles, straight-line code.
//   * warp 4     issues the dependent MMA stages (forward layers, dH) of both slots back to back, one commit.
//   * warps 5,6  issue the weight-gradient MMAs (8 x M=128 x N=80 per tile) of slot 0 / slot 1 into their own
//                TMEM accumulators (summed at the end), off everybody's critical path.
//   * warp 7     (features mode) keeps fp32 features in flight with cp.async.bulk.
//   * the bias and the all-ones feature ride inside the MMA: weight-image row in_l holds b_l and a 1.0 that
//     re-creates the ones column in the next layer's input, so a hidden-layer epilogue is tcgen05.ld ->
//     cvt.rn.relu.bf16x2 -> 16-byte stores; the ReLU adjoint mask is one HSET2 + LOP3 per pair.
//   * shared memory per tile shrinks from 48 to 44 KB (A_0 holds ceil((c_in+1)/8) slabs; K padding reads the
//     next buffer's finite values against zero weight rows): two 2-tile CTAs per SM, 4 tiles in flight.

constexpr int PP_ROWS = 128;                 // row threads (= TILE)
constexpr int PP_THREADS = PP_ROWS + 128;    // + MMA issuer warp, two dW issuer warps, TMA producer warp
constexpr int PP_MAX_STAGES = 2 * MAXL;      // fwd 0..L-1, dH 1..L-1, dW

template <int HP>
struct PpLayout {
    static constexpr int HSL = HP / 8;
    __host__ __device__ static int a0s(int c_in) { return (c_in + 8) >> 3; }   // slabs holding features 0..c_in (ones column incl.)
    __host__ __device__ static int ndw(int L) { return (L - 1) * HP + 16; }
    __host__ __device__ static int a_off(int l, int A0S) { return l == 0 ? 0 : (A0S + (l - 1) * HSL) * SLAB; }
    __host__ __device__ static int dz_off(int l, int L, int A0S) { return (A0S + (L - 1) * HSL + l * HSL) * SLAB; }
    // grad: A_0 .. A_{L-1} | dZ_0 .. dZ_{L-2} | dZ_{L-1} (ONE slab: its features 8..15 read whatever follows, finite,
    // against zero weight rows).  forward only: the A buffers.
    __host__ __device__ static int slot_bytes(int L, int A0S, bool grad)
    {
        return grad ? (A0S + 2 * (L - 1) * HSL + 1) * SLAB : (A0S + (L - 1) * HSL) * SLAB;
    }
    __host__ __device__ static int np(int l, int L) { return l < L - 1 ? HP : 16; }
    __host__ __device__ static int kp(int l, int K0P) { return l == 0 ? K0P : HP; }
    __host__ __device__ static int w_off(int l, int L, int K0P)
    {
        int o = 0;
        for (int i = 0; i < l; ++i) o += np(i, L) * kp(i, K0P) * 2;
        return o;
    }
    __host__ __device__ static int wimg_bytes(int L, int K0P) { return w_off(L, L, K0P) + MAXL * HP * 4; }
    __host__ __device__ static int stage_bytes(int c_in, int K0P) { return (TILE * c_in * 4 + 32 + K0P * 4 + 15) / 16 * 16; }
    __host__ __device__ static uint32_t tmem_cols(int L, int nslot, bool grad)
    {
        const int need = nslot * HP + (grad ? nslot * ndw(L) : 0);
        return need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512)));
    }
    static constexpr int MISC_BYTES = 128 * 4 + 16 * 8 + 16;  // scan scratch (two slots), mbarriers, TMEM slot
    __host__ __device__ static int ray_scratch_bytes(bool grad) { return grad ? 0 : TILE * 3 * 4 * 2; }   // colour + target per ray
    __host__ __device__ static size_t total(int L, int K0P, int c_in, int nslot, bool rays, bool grad)
    {
        const int A0S = a0s(c_in);
        size_t t = (size_t)nslot * slot_bytes(L, A0S, grad) + wimg_bytes(L, K0P) + (rays ? 0 : stage_bytes(c_in, K0P)) + nslot * ray_scratch_bytes(grad) +
                   MISC_BYTES + nslot * PP_MAX_STAGES * 48;
        // the dW MMA reads 16 slabs (M = 128 feature rows) from a slot's base whatever the real feature count
        const size_t need = (size_t)(nslot - 1) * slot_bytes(L, A0S, grad) + 16 * SLAB;
        if (grad && t < need) t = need;
        return (t + 15) / 16 * 16;
    }
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bar_rows() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// hand a buffer over: every lane has fenced its own writes; one arrival per warp (the barrier counts warps)
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane)
{
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
#ifdef LNB_TC_CLK
#define PCLK_DECL const bool pclk_on = (threadIdx.x & 127) == 0; long long pclk_t = clock64(); long long pclk[24] = {0}
#define PCLK(i) do { if (pclk_on) { const long long t_ = clock64(); pclk[i] += t_ - pclk_t; pclk_t = t_; } } while (0)
#define PCLK_OUT(first, n) do { if (p.dbg) for (int i_ = (first); i_ < (first) + (n); ++i_) p.dbg[blockIdx.x * 24 + i_] = (float)pclk[i_]; } while (0)
#else
#define PCLK_DECL
#define PCLK(i)
#define PCLK_OUT(first, n)
#endif

template <bool RAYS, int HP, int NSLOT>
__global__ void __launch_bounds__(PP_THREADS, 2) fused_tc_kernel(const TcParams p)
{
#ifdef LNB_TC_CLK
    unsigned long long gt_start;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_start));
#endif
    using LY = PpLayout<HP>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.L, K0P = p.K0P, c_in = p.dims[0], S = p.S;
    const bool grad = p.want_grad != 0;
    const int A0S = LY::a0s(c_in);
    const int slot_bytes = LY::slot_bytes(L, A0S, grad);
    uint8_t *const Wbase = smem + NSLOT * slot_bytes;
    const float *const bias_s = reinterpret_cast<const float *>(Wbase + LY::w_off(L, L, K0P));
    float *const stage = reinterpret_cast<float *>(Wbase + LY::wimg_bytes(L, K0P));
    const int stage_sz = RAYS ? 0 : LY::stage_bytes(c_in, K0P);
    float *const ray_scr = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(stage) + stage_sz);          // forward only
    float *const scan_scr = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(ray_scr) + NSLOT * LY::ray_scratch_bytes(grad));
    float *const red_s = scan_scr + 112;
    uint64_t *const bar_p = reinterpret_cast<uint64_t *>(scan_scr + 128);
    // barriers: 0 in (row warps -> MMA issuer), 1 out (commit -> rows), 2 dwi (rows -> dW issuers), 3 dwo (their commits -> rows),
    //           4 weights, 5 x_full, 6 x_empty
    const uint32_t bar0 = smem_u32(bar_p);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    uint32_t *const tmem_slot = reinterpret_cast<uint32_t *>(bar_p + 16);
    struct StageRec { uint64_t a, b; uint32_t inc_a, inc_b, idesc, dcol; uint32_t count, first_acc, pad0, pad1; };
    StageRec *const prog = reinterpret_cast<StageRec *>(reinterpret_cast<uint8_t *>(bar_p) + 16 * 8 + 16);

    // ---- one-time setup
    {
        const int zero_end = (int)(reinterpret_cast<uint8_t *>(scan_scr) - smem) + LY::MISC_BYTES;
        for (int o = tid * 16; o < zero_end; o += PP_THREADS * 16) *reinterpret_cast<uint4 *>(smem + o) = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    if (tid == 0) {
        mbar_init(BAR(0), PP_ROWS / 32);
        mbar_init(BAR(1), 1);
        mbar_init(BAR(2), PP_ROWS / 32);
        mbar_init(BAR(3), NSLOT);
        mbar_init(BAR(4), 1);
        mbar_init(BAR(5), 1);
        mbar_init(BAR(6), PP_ROWS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < NSLOT; ++s) {
            uint8_t *sb = smem + s * slot_bytes;
            StageRec *pr = prog + s * PP_MAX_STAGES;
            const uint32_t dres = (uint32_t)(s * HP);
            for (int l = 0; l < L; ++l) {          // D[128 x Np] = A_l[128 x Kp] * W_l   (A, B K-major)
                const int Np = LY::np(l, L), Kp = LY::kp(l, K0P);
                const uint32_t a0 = smem_u32(sb + LY::a_off(l, A0S)), b0 = smem_u32(Wbase + LY::w_off(l, L, K0P));
                pr[l] = StageRec{smem_desc(a0, SLAB, 128), smem_desc(b0, Np * 16, 128), (uint32_t)(2 * SLAB) >> 4, (uint32_t)(2 * Np * 16) >> 4,
                                 instr_desc(128, Np, 0, 0), dres, (uint32_t)(Kp / 16), 0u, 0u, 0u};
            }
            if (grad) {
                for (int l = 1; l < L; ++l) {      // dH_l[128 x Kp] = dZ_l[128 x Np] * W_l^T (same W bytes, MN-major); stage index 2L-1-l
                    const int Np = LY::np(l, L), Kp = LY::kp(l, K0P);
                    const uint32_t a0 = smem_u32(sb + LY::dz_off(l, L, A0S)), b0 = smem_u32(Wbase + LY::w_off(l, L, K0P));
                    pr[2 * L - 1 - l] = StageRec{smem_desc(a0, SLAB, 128), smem_desc(b0, 128, Np * 16), (uint32_t)(2 * SLAB) >> 4, 256u >> 4,
                                                 instr_desc(128, Kp, 0, 1), dres, (uint32_t)(Np / 16), 0u, 0u, 0u};
                }
                const uint32_t a0 = smem_u32(sb), b0 = smem_u32(sb + LY::dz_off(0, L, A0S));   // dW: [all A]^T [all dZ], K = 128 samples
                pr[2 * L - 1] = StageRec{smem_desc(a0, 128, SLAB), smem_desc(b0, 128, SLAB), 256u >> 4, 256u >> 4,
                                         instr_desc(128, LY::ndw(L), 1, 1), (uint32_t)(NSLOT * HP + s * LY::ndw(L)), (uint32_t)(TILE / 16), 2u, 0u, 0u};
            }
        }
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), LY::tmem_cols(L, NSLOT, grad));
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    asm volatile("griddepcontrol.wait;" ::: "memory");              // from here on we read what the previous kernel wrote
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    // work units: groups of NSLOT consecutive tiles; group g of this CTA = blockIdx.x + i * gridDim.x
    const int n_groups = (p.n_tiles + NSLOT - 1) / NSLOT;
    const int n_crit = grad ? 2 * L - 1 : L;

    // TMA source of a tile's features (16 B aligned start, `lead` floats in front of the tile)
    auto x_src = [&](int tile, int &lead, uint32_t &bytes) -> const void * {
        const long long row0 = (long long)tile * p.rows_per_tile;
        long long rem = p.N - row0;
        const int valid = rem < p.rows_per_tile ? (int)rem : p.rows_per_tile;
        const uintptr_t a = reinterpret_cast<uintptr_t>(p.X + row0 * c_in);
        const uintptr_t a16 = a & ~(uintptr_t)15;
        lead = (int)((a - a16) >> 2);
        bytes = (uint32_t)(((a - a16) + (uintptr_t)valid * c_in * 4 + 15) & ~(uintptr_t)15);
        return reinterpret_cast<const void *>(a16);
    };
    auto issue_stage = [&](const StageRec *rec, uint32_t first_acc) {
        const uint4 r0 = *reinterpret_cast<const uint4 *>(rec);
        const uint4 r1 = *(reinterpret_cast<const uint4 *>(rec) + 1);
        const uint4 r2 = *(reinterpret_cast<const uint4 *>(rec) + 2);
        uint32_t alo = r0.x, blo = r0.z, acc = first_acc;
        const uint32_t ahi = r0.y, bhi = r0.w, d = tmem + r1.w;
        for (uint32_t k = 0; k < r2.x; ++k) {
            umma_bf16(d, ((uint64_t)ahi << 32) | alo, ((uint64_t)bhi << 32) | blo, r1.z, acc);
            alo += r1.x; blo += r1.y; acc = 1u;
        }
    };

    if (warp >= 4) {
        // =========================== the single-thread roles ===========================
        if (lane == 0) {
            if (warp == 4) {
                // ---- MMA issuer: per group and stage, both slots' MMAs back to back, one commit
                mbar_wait(BAR(4), 0);   // weights have landed
                uint32_t par = 0;
                for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
                    // (a group's missing second tile is run as a dead tile: finite stale operands, zero adjoints)
                    for (int st = 0; st < n_crit; ++st) {
                        mbar_wait(BAR(0), par);
                        par ^= 1;
                        tc_fence_after();
                        issue_stage(prog + st, 0u);
                        if (NSLOT == 2) issue_stage(prog + PP_MAX_STAGES + st, 0u);
                        umma_commit(BAR(1));
                    }
                }
            } else if (warp < 4 + 1 + NSLOT) {
                if (grad) {
                    // ---- dW issuer of slot s: the concatenated H^T dZ product of the slot's tile, accumulated over all groups
                    const int s = warp - 5;
                    mbar_wait(BAR(4), 0);
                    uint32_t par = 0, started = 0;
                    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
                        mbar_wait(BAR(2), par);
                        par ^= 1;
                        tc_fence_after();
                        issue_stage(prog + s * PP_MAX_STAGES + 2 * L - 1, started);
                        started = 1;
                        umma_commit(BAR(3));
                    }
                }
            } else if (warp == 7) {
                // ---- TMA producer: the weight image once, then every tile's features in the order the row threads
                // convert them; one stage buffer, refilled as soon as it has been read
                const uint32_t wb = (uint32_t)LY::wimg_bytes(L, K0P);
                mbar_expect_tx(BAR(4), wb);
                bulk_g2s(smem_u32(Wbase), p.wimg, wb, BAR(4));
                if (!RAYS) {
                    uint32_t n_loads = 0;
                    for (int g = blockIdx.x; g < n_groups; g += gridDim.x)
                        for (int s = 0; s < NSLOT; ++s) {
                            const int tile = g * NSLOT + s;
                            if (tile >= p.n_tiles) break;
                            if (n_loads > 0) mbar_wait(BAR(6), (n_loads - 1) & 1u);
                            int lead; uint32_t bytes;
                            const void *src = x_src(tile, lead, bytes);
                            mbar_expect_tx(BAR(5), bytes);
                            bulk_g2s(smem_u32(stage), src, bytes, BAR(5));
                            ++n_loads;
                        }
                }
            }
        }
    } else {
        // ================================= row threads =================================
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        float loss_acc = 0.0f;
        uint32_t n_conv = 0, out_par = 0, dwo_par = 0;
        bool dw_pending = false;
        const int smp = tid % S, ray_l = tid / S;          // this thread's sample within its ray: the same in every tile
        const int full_live = (p.rows_per_tile / S) * S;
        auto row_ptr = [&](uint8_t *buf, int slab) { return reinterpret_cast<uint4 *>(buf + slab * SLAB + tid * 16); };
        auto wait_out = [&]() { mbar_wait(BAR(1), out_par); out_par ^= 1; tc_fence_after(); };
        mbar_wait(BAR(4), 0);   // biases are read from the image
        PCLK_DECL;
        PCLK(22);
        for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
            long long row0[NSLOT];
            int valid[NSLOT];
            bool live[NSLOT];
            float dist[NSLOT], tg[NSLOT][3];
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) {
                const int tile = g * NSLOT + s;
                const bool has = tile < p.n_tiles;                  // a missing second tile is a dead tile: no loads, zero adjoints
                row0[s] = (long long)tile * p.rows_per_tile;
                const bool last = tile == p.n_tiles - 1;
                valid[s] = !has ? 0 : (last ? (int)(p.N - row0[s]) : p.rows_per_tile);
                live[s] = has && (last ? tid < (valid[s] / S) * S : tid < full_live);
                dist[s] = 0.f; tg[s][0] = tg[s][1] = tg[s][2] = 0.f;
                if (p.head == LNB_HEAD_NERF && live[s]) {
                    if (!RAYS) dist[s] = __ldg(p.dists + row0[s] + tid);
                    if (p.target && smp == 0) {
                        const float *t3 = p.target + (row0[s] / S + ray_l) * 3;
                        tg[s][0] = __ldg(t3); tg[s][1] = __ldg(t3 + 1); tg[s][2] = __ldg(t3 + 2);
                    }
                }
            }
            // ---------------- phase 0: features of both tiles -> A_0 ----------------
            if (RAYS) {
                // pts = o + d t (train_nerf.py:289-299), PE (pos_encoding.py:38-70), dist = t[s+1] - t[s], last 1e8
                // (train_nerf.py:306-311); the bf16 rows are built in registers and leave as 16-byte stores
                float x[NSLOT][3], sn[NSLOT][3], cs[NSLOT][3];
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    x[s][0] = x[s][1] = x[s][2] = 0.f;
                    if (live[s]) {
                        const long long ray = row0[s] / S + ray_l, smpl = row0[s] + tid;
                        if (p.ray_f64) {
                            const double *o = reinterpret_cast<const double *>(p.rays_o) + ray * 3;
                            const double *d = reinterpret_cast<const double *>(p.rays_d) + ray * 3;
                            const double *tv = reinterpret_cast<const double *>(p.tvals) + smpl;
                            const double tt = __ldg(tv);
                            dist[s] = smp + 1 < S ? (float)(__ldg(tv + 1) - tt) : 1e8f;
#pragma unroll
                            for (int c = 0; c < 3; ++c) x[s][c] = (float)(__ldg(o + c) + __ldg(d + c) * tt);
                        } else {
                            const float *o = reinterpret_cast<const float *>(p.rays_o) + ray * 3;
                            const float *d = reinterpret_cast<const float *>(p.rays_d) + ray * 3;
                            const float *tv = reinterpret_cast<const float *>(p.tvals) + smpl;
                            const float tt = __ldg(tv);
                            dist[s] = smp + 1 < S ? __ldg(tv + 1) - tt : 1e8f;
#pragma unroll
                            for (int c = 0; c < 3; ++c) x[s][c] = fmaf(__ldg(d + c), tt, __ldg(o + c));
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c) pe_sincos(x[s][c], &sn[s][c], &cs[s][c]);
                }
                PCLK(8);
                if (dw_pending) { mbar_wait(BAR(3), dwo_par); dwo_par ^= 1; dw_pending = false; }   // the previous group's dW MMAs read A_0
                PCLK(0);
                // feature f = 3 * slot + coord: x0 x1 | x2 s0 | s1 s2 | c0 c1 | c2 s0' | ... as bf16 pairs (pair j = features 2j, 2j+1); band i
                // fills pairs 1+3i .. 3+3i and leaves its last cosine pending; the ones column (feature 3 + 6E, odd) closes the pending
                // pair.  Four pairs make a 16-byte slab entry, stored as soon as it is complete.
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    uint8_t *const a0 = smem + s * slot_bytes;
                    uint32_t q4[4] = {0u, 0u, 0u, 0u};
                    float pend = x[s][2];
                    bool closed = false;
                    q4[0] = pack_bf16(x[s][0], x[s][1]);
                    auto put = [&](int j, uint32_t v) {            // j is a compile-time constant at every call site
                        q4[j & 3] = v;
                        if ((j & 3) == 3) { if ((j >> 2) < A0S) *row_ptr(a0, j >> 2) = make_uint4(q4[0], q4[1], q4[2], q4[3]); q4[0] = q4[1] = q4[2] = q4[3] = 0u; }
                    };
#pragma unroll
                    for (int i = 0; i < 10; ++i) {
                        const bool on = i < p.pe_bands;
                        const bool close = !closed && !on;
                        put(1 + 3 * i, on ? pack_bf16(pend, sn[s][0]) : (close ? pack_bf16(pend, 1.0f) : 0u));
                        put(2 + 3 * i, on ? pack_bf16(sn[s][1], sn[s][2]) : 0u);
                        put(3 + 3 * i, on ? pack_bf16(cs[s][0], cs[s][1]) : 0u);
                        closed = closed || close;
                        pend = cs[s][2];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float s2 = 2.0f * sn[s][c] * cs[s][c], c2 = fmaf(-2.0f * sn[s][c], sn[s][c], 1.0f);
                            sn[s][c] = s2; cs[s][c] = c2;
                        }
                    }
                    put(31, closed ? 0u : pack_bf16(pend, 1.0f));
                }
            } else {
                // wait for the TMA, convert this thread's row to bf16 slabs.  Columns beyond c_in read the following floats
                // of `stage` (finite: next row / zeroed slack) and meet zero weight rows.
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    uint8_t *const a0 = smem + s * slot_bytes;
                    if (g * NSLOT + s >= p.n_tiles) {              // dead tile: keep the stale (finite) features; its adjoints are zero
                        if (dw_pending) { mbar_wait(BAR(3), dwo_par); dwo_par ^= 1; dw_pending = false; }
                        continue;
                    }
                    int lead; uint32_t bytes;
                    (void)x_src(g * NSLOT + s, lead, bytes);
                    PCLK(8);
                    mbar_wait(BAR(5), n_conv & 1u);
                    ++n_conv;
                    PCLK(21);
                    if (dw_pending) { mbar_wait(BAR(3), dwo_par); dwo_par ^= 1; dw_pending = false; }
                    PCLK(0);
                    const float *xr = stage + lead + tid * c_in;
                    if (tid < valid[s]) {
                        for (int c8 = 0; c8 < A0S; ++c8) {
                            float f[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = xr[c8 * 8 + j];
                            *row_ptr(a0, c8) = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                        }
                    } else {
                        for (int c8 = 0; c8 < A0S; ++c8) *row_ptr(a0, c8) = make_uint4(0, 0, 0, 0);
                    }
                    warp_arrive(BAR(6), lane);   // this warp is done with the stage buffer
                    // the all-ones feature (bias, bias gradient); same thread, same bytes as the vector store above
                    reinterpret_cast<__nv_bfloat16 *>(a0)[(c_in >> 3) * (TILE * 8) + tid * 8 + (c_in & 7)] = __float2bfloat16_rn(1.0f);
                }
            }
            fence_async_smem();
            warp_arrive(BAR(0), lane);
            PCLK(8);
            // ---------------- forward epilogues: layer l-1 -> A_l (bias and ones column came out of the MMA) ----------------
            for (int l = 1; l < L; ++l) {
                wait_out();
                PCLK(1);
                uint32_t v[NSLOT][HP / 16][16];
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
#pragma unroll
                    for (int c16 = 0; c16 < HP / 16; ++c16) tmem_ld16(tmem + lane_base + (uint32_t)(s * HP + c16 * 16), v[s][c16]);
                tmem_ld_wait();
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    uint8_t *const an = smem + s * slot_bytes + LY::a_off(l, A0S);
#pragma unroll
                    for (int c16 = 0; c16 < HP / 16; ++c16) {
                        uint32_t o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = pack_relu_bf16(__uint_as_float(v[s][c16][2 * j]), __uint_as_float(v[s][c16][2 * j + 1]));
                        *row_ptr(an, c16 * 2) = make_uint4(o[0], o[1], o[2], o[3]);
                        *row_ptr(an, c16 * 2 + 1) = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
                fence_async_smem();
                tc_fence_before();
                warp_arrive(BAR(0), lane);
                PCLK(9);
            }
            // ---------------- head, loss, adjoint of the head's pre-activation (unit seed) ----------------
            wait_out();
            PCLK(3);
            float hz[NSLOT][4];
            {
                uint32_t v[NSLOT][16];
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) tmem_ld16(tmem + lane_base + (uint32_t)(s * HP), v[s]);
                tmem_ld_wait();
                const float *bl = bias_s + (L - 1) * HP;
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
#pragma unroll
                    for (int j = 0; j < 4; ++j) hz[s][j] = __uint_as_float(v[s][j]) + bl[j];
            }
            float dz[NSLOT][4];
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) dz[s][0] = dz[s][1] = dz[s][2] = dz[s][3] = 0.f;
            if (p.head == LNB_HEAD_SIGMOID) {
                // mlp_fit: row r <-> target row (scripts/mlp_fit.py:121-145)
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
                    if (tid < valid[s] && row0[s] + tid < p.R) {
                        const float *t3 = p.target + (row0[s] + tid) * p.Wt;
                        for (int c = 0; c < p.Wt && c < 4; ++c) {
                            float y = sigmoid_f(hz[s][c]);
                            float d = y - __ldg(t3 + c);
                            loss_acc = fmaf(d, d, loss_acc);
                            dz[s][c] = 2.0f * d * (y * (1.0f - y));
                        }
                    }
            } else {
                // Compositing with one thread per sample, both tiles side by side: segmented warp-shuffle scans inside each
                // warp, carries across the 4 warps through shared memory.  scripts/nerf.py:176-288 and its reverse (App. B).
                float cr[NSLOT], cg[NSLOT], cb[NSLOT], sg[NSLOT], e[NSLOT], a[NSLOT], qv[NSLOT], pr[NSLOT];
                float *color_s[NSLOT], *tgt_s[NSLOT], *tailp[NSLOT], *headq[NSLOT], *headA[NSLOT], *headB[NSLOT];
                int *tail_s[NSLOT];
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    color_s[s] = grad ? reinterpret_cast<float *>(smem + s * slot_bytes + LY::dz_off(0, L, A0S)) : ray_scr + s * (TILE * 6);   // dZ_0 is dead here
                    tgt_s[s] = color_s[s] + TILE * 3;
                    tailp[s] = scan_scr + s * 56; tail_s[s] = reinterpret_cast<int *>(tailp[s] + 4);
                    headq[s] = tailp[s] + 8; headA[s] = tailp[s] + 13; headB[s] = tailp[s] + 18;
                    cr[s] = sigmoid_f(hz[s][0]); cg[s] = sigmoid_f(hz[s][1]); cb[s] = sigmoid_f(hz[s][2]);
                    sg[s] = fmaxf(hz[s][3], 0.0f);
                    e[s] = __expf((0.0f - sg[s]) * dist[s]);
                    a[s] = 1.0f - e[s];
                    qv[s] = live[s] ? (1.0f - a[s]) + 1e-10f : 1.0f;
                    pr[s] = qv[s];                                   // segmented inclusive product
                }
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s) {
                        float o = __shfl_up_sync(0xffffffffu, pr[s], d);
                        if (lane >= d && smp >= d) pr[s] *= o;
                    }
                }
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    if (lane == 31) { tailp[s][warp] = pr[s]; tail_s[s][warp] = smp; }
                    if (lane == 0) headq[s][warp] = qv[s];
                    if (live[s] && smp == 0) {
                        color_s[s][ray_l * 3] = 0.f; color_s[s][ray_l * 3 + 1] = 0.f; color_s[s][ray_l * 3 + 2] = 0.f;
                        tgt_s[s][ray_l * 3] = tg[s][0]; tgt_s[s][ray_l * 3 + 1] = tg[s][1]; tgt_s[s][ray_l * 3 + 2] = tg[s][2];
                    }
                }
                bar_rows();
                float carry[NSLOT], Cpre[NSLOT], T[NSLOT], wgt[NSLOT];
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    carry[s] = 1.0f;                                 // product of this ray's samples in earlier warps
                    if (smp > lane) {
                        for (int w2 = warp - 1; w2 >= 0; --w2) {
                            carry[s] *= tailp[s][w2];
                            if (tail_s[s][w2] < 32) break;           // that warp's last segment started inside it
                        }
                    }
                    Cpre[s] = pr[s] * carry[s];                      // true inclusive product prod_{k<=s} q_k
                    T[s] = (smp == 0) ? 1.0f : Cpre[s];
                    wgt[s] = a[s] * T[s];
                }
                {   // colour: segmented inclusive sums, one shared-memory atomic per (warp, ray) segment
                    float s0[NSLOT], s1[NSLOT], s2[NSLOT];
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s) {
                        s0[s] = live[s] ? wgt[s] * cr[s] : 0.f; s1[s] = live[s] ? wgt[s] * cg[s] : 0.f; s2[s] = live[s] ? wgt[s] * cb[s] : 0.f;
                    }
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                        for (int s = 0; s < NSLOT; ++s) {
                            float o0 = __shfl_up_sync(0xffffffffu, s0[s], d), o1 = __shfl_up_sync(0xffffffffu, s1[s], d), o2 = __shfl_up_sync(0xffffffffu, s2[s], d);
                            if (lane >= d && smp >= d) { s0[s] += o0; s1[s] += o1; s2[s] += o2; }
                        }
                    }
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s)
                        if (live[s] && (lane == 31 || smp == S - 1)) {
                            atomicAdd(color_s[s] + ray_l * 3, s0[s]); atomicAdd(color_s[s] + ray_l * 3 + 1, s1[s]); atomicAdd(color_s[s] + ray_l * 3 + 2, s2[s]);
                        }
                }
                bar_rows();
                float dc0[NSLOT], dc1[NSLOT], dc2[NSLOT];
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    dc0[s] = dc1[s] = dc2[s] = 0.f;
                    if (live[s]) {
                        const float c0 = color_s[s][ray_l * 3], c1 = color_s[s][ray_l * 3 + 1], c2 = color_s[s][ray_l * 3 + 2];
                        if (smp == 0 && p.color) {
                            float *co = p.color + (row0[s] / S + ray_l) * 3;
                            co[0] = c0; co[1] = c1; co[2] = c2;
                        }
                        if (p.target) {
                            const float d0 = c0 - tgt_s[s][ray_l * 3], d1 = c1 - tgt_s[s][ray_l * 3 + 1], d2 = c2 - tgt_s[s][ray_l * 3 + 2];
                            if (smp == 0) loss_acc += d0 * d0 + d1 * d1 + d2 * d2;
                            dc0[s] = 2.0f * d0; dc1[s] = 2.0f * d1; dc2[s] = 2.0f * d2;
                        }
                    }
                }
                if (grad && p.target) {
                    // G_s = dT_s + q_{s+1} G_{s+1}: suffix scan of affine maps; B = 0 at a ray's last sample
                    float d_w[NSLOT], Aa[NSLOT], Bb[NSLOT];
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s) {
                        d_w[s] = cr[s] * dc0[s] + cg[s] * dc1[s] + cb[s] * dc2[s];
                        const float dT = (smp == 0 || !live[s]) ? 0.0f : d_w[s] * a[s];
                        float qn = __shfl_down_sync(0xffffffffu, qv[s], 1);
                        if (lane == 31) qn = warp < 3 ? headq[s][warp + 1] : 0.0f;
                        Aa[s] = dT; Bb[s] = (live[s] && smp + 1 < S) ? qn : 0.0f;
                    }
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                        for (int s = 0; s < NSLOT; ++s) {
                            float A2 = __shfl_down_sync(0xffffffffu, Aa[s], d);
                            float B2 = __shfl_down_sync(0xffffffffu, Bb[s], d);
                            if (lane + d < 32) { Aa[s] = fmaf(Bb[s], A2, Aa[s]); Bb[s] = Bb[s] * B2; }
                        }
                    }
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s)
                        if (lane == 0) { headA[s][warp] = Aa[s]; headB[s][warp] = Bb[s]; }
                    bar_rows();
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s) {
                        float Gn = 0.0f;                             // G at lane 0 of the next warp
                        for (int w2 = 3; w2 > warp; --w2) Gn = fmaf(headB[s][w2], Gn, headA[s][w2]);
                        const float Gv = fmaf(Bb[s], Gn, Aa[s]);
                        float Cm1 = __shfl_up_sync(0xffffffffu, Cpre[s], 1);
                        if (lane == 0) Cm1 = carry[s];
                        if (smp == 0) Cm1 = 1.0f;
                        const float d_alpha = d_w[s] * T[s] - Cm1 * Gv;
                        if (live[s]) {
                            dz[s][0] = (wgt[s] * dc0[s]) * (cr[s] * (1.0f - cr[s]));
                            dz[s][1] = (wgt[s] * dc1[s]) * (cg[s] * (1.0f - cg[s]));
                            dz[s][2] = (wgt[s] * dc2[s]) * (cb[s] * (1.0f - cb[s]));
                            dz[s][3] = sg[s] > 0.0f ? d_alpha * e[s] * dist[s] : 0.0f;
                        }
                    }
                }
            }
            if (!grad) {
                // forward only: the next group's first MMA (ordered by the arrival after its A_0) may overwrite TMEM
                tc_fence_before();
                continue;
            }
            // dZ_{L-1}: 4 live features; features 4..7 zero, 8..15 belong to whatever follows (zero weight rows)
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
                *row_ptr(smem + s * slot_bytes + LY::dz_off(L - 1, L, A0S), 0) = make_uint4(pack_bf16(dz[s][0], dz[s][1]), pack_bf16(dz[s][2], dz[s][3]), 0u, 0u);
            fence_async_smem();
            tc_fence_before();
            warp_arrive(BAR(0), lane);
            PCLK(11);
            // ---------------- backward epilogues: dH_l -> dZ_{l-1} = dH_l * [A_l > 0] ----------------
            for (int l = L - 1; l >= 1; --l) {
                wait_out();
                PCLK(4);
                uint32_t v[NSLOT][HP / 16][16];
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
#pragma unroll
                    for (int c16 = 0; c16 < HP / 16; ++c16) tmem_ld16(tmem + lane_base + (uint32_t)(s * HP + c16 * 16), v[s][c16]);
                tmem_ld_wait();
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    uint8_t *const al = smem + s * slot_bytes + LY::a_off(l, A0S);
                    uint8_t *const dzn = smem + s * slot_bytes + LY::dz_off(l - 1, L, A0S);
#pragma unroll
                    for (int c8 = 0; c8 < HP / 8; ++c8) {
                        const uint4 hm = *row_ptr(al, c8);
                        const uint32_t hw[4] = {hm.x, hm.y, hm.z, hm.w};
                        uint32_t o[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int ee = (c8 & 1) * 8 + 2 * j;
                            o[j] = pack_bf16(__uint_as_float(v[s][c8 >> 1][ee]), __uint_as_float(v[s][c8 >> 1][ee + 1])) & gt0_mask_bf16x2(hw[j]);
                        }
                        *row_ptr(dzn, c8) = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                }
                fence_async_smem();
                tc_fence_before();
                if (l > 1) warp_arrive(BAR(0), lane);
                else { warp_arrive(BAR(2), lane); dw_pending = true; }
                PCLK(12);
            }
        }
        // ---- this CTA's partials: loss, then per layer the valid (in_l+1) x out_l block of the dW accumulators
        if (dw_pending) mbar_wait(BAR(3), dwo_par);
        tc_fence_after();
        PCLK(23);
#ifdef LNB_TC_CLK
        if (tid == 0) { PCLK_OUT(0, 17); PCLK_OUT(21, 3); }
#endif
        float *part = p.part + (size_t)blockIdx.x * p.part_stride;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, d);
        if (lane == 0) red_s[warp] = loss_acc;
        bar_rows();
        if (tid == 0) part[0] = red_s[0] + red_s[1] + red_s[2] + red_s[3];
        if (tid == 0 && blockIdx.x == 0 && p.t_dev) p.t_dev[0] += 1; // read by the kernels that follow in the stream
        if (grad) {
            // accumulator row (TMEM lane) = feature index over the concatenated A buffers; thread tid owns row tid
            const bool two = NSLOT == 2;   // slot 1's accumulator has been written by every group (zeros for a dead tile)
            for (int l = 0; l < L; ++l) {
                const int rowbase = LY::a_off(l, A0S) / SLAB * 8;
                const int in_l = p.dims[l], out_l = p.dims[l + 1], Np = LY::np(l, L);
                const int row = tid - rowbase;
                const int colbase = NSLOT * HP + l * HP;
                float *o = part + p.part_off[l];
#pragma unroll
                for (int c16 = 0; c16 < HP / 16; ++c16) {
                    if (c16 * 16 >= Np) break;
                    uint32_t v[16], v1[16];
                    tmem_ld16(tmem + lane_base + (uint32_t)(colbase + c16 * 16), v);
                    if (two) tmem_ld16(tmem + lane_base + (uint32_t)(colbase + LY::ndw(L) + c16 * 16), v1);
                    tmem_ld_wait();
                    if (row >= 0 && row <= in_l) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            int col = c16 * 16 + j;
                            if (col < out_l) o[row * out_l + col] = __uint_as_float(v[j]) + (two ? __uint_as_float(v1[j]) : 0.0f);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, LY::tmem_cols(L, NSLOT, grad));
#ifdef LNB_TC_CLK
    if (p.dbg && tid == 0) {
        unsigned long long gt_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_end));
        unsigned long long *tt = reinterpret_cast<unsigned long long *>(p.dbg + (size_t)gridDim.x * 24);
        tt[blockIdx.x * 2] = gt_start; tt[blockIdx.x * 2 + 1] = gt_end;
    }
#endif
}
