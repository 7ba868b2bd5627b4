// spmm_variants.cuh -- adjoint-SpMM kernels that were built, validated on the GPU (agreement with the shipped kernel to 1e-15
// relative, tests of commit "SpMM experiments") and measured in round 2, and that LOST to the shipped kernels on configs[2]:
// the row-stream DFMA kernel (292 us) and the FIRST version of the TMA gather4-fed DMMA kernel (290 us; its second version,
// with the row indices passed through REDUX into uniform registers, is in csrc/spmm.cu as k_spmm_adj_tma: 238 us).  See
// profiles/r02_spmm_attempts.md for the ncu numbers and what each one is bound by.  Kept as a record; NOT compiled.
#if 0
// ===================================================== kernels ======================================================
// Adjoint SpMM as a row STREAM (Float64, rows mode, one stripe width W for the whole matrix, a panel of <= 32 right-hand
// sides): with a uniform width, stored row p of the matrix is desc[p] / val[p*W, (p+1)*W) and stripe l owns columns
// [l*W, (l+1)*W), so a warp's work is a flat sequence of rows cut into stripes by meta[].pos.  Lane (h, cc) = (lane >> 4,
// lane & 15) owns right-hand sides 2cc, 2cc+1 and, of every PAIR of rows, row h: one 128-bit load per lane fetches two whole
// gathered X rows per warp instruction (512 contiguous-by-row bytes: two L1 wavefronts per row, the minimum -- the DMMA
// fragment layout needs eight), the row's W values come from shared memory as half-warp-uniform LDS.128, and W x 2 DFMAs
// per lane follow.  What is in flight, without a shared-memory ring for X: the X rows of the NEXT chunk of 16 rows sit in a
// rolling register ring (slot j is reloaded right after pair j is consumed), the val chunks of the next two chunks arrive
// through cp.async, the x indices three chunks ahead in a register.  Units of ~8 stripes are dealt round-robin to the warps
// (all warps work in one sliding window of X: L2 reuse of the gathered rows); the per-stripe result is the sum of the two
// half-warps (16 shuffles) and is stored as 256-byte row segments.  ~16 instructions per stored row against ~45 in
// k_spmm_adj; DFMA issue, not the L1 data pipe, is the next limit after HBM.
template <int W, bool FULL>
__global__ void __launch_bounds__(256, 2) k_spmm_adj_stream(const StripeMeta *__restrict__ meta, const int *__restrict__ desc,
                                                            const double *__restrict__ val, const double *__restrict__ X, const unsigned ldx_bytes,
                                                            double *__restrict__ Y, const long long ldy, const int L, const int nunits,
                                                            const double ratio, const int k, const int kb, const double alpha, const double beta)
{
    constexpr int CR = 16, NP = CR / 2, NV = 3, HW = W / 2;
    constexpr int DEAD = INT_MIN; // seg_end of a unit whose stripes are all stored: nothing is multiplied until the next unit begins
    __shared__ __align__(16) double vs_all[8][NV][CR * W];
    const int lane = threadIdx.x & 31, h = lane >> 4, cc = lane & 15;
    const int nwarps = (int)gridDim.x * 8, wid = (int)blockIdx.x * 8 + (int)(threadIdx.x >> 5);
    if (wid >= nunits) return;
    double(*vs)[CR * W] = vs_all[threadIdx.x >> 5];
    const bool colok = FULL || kb + 2 * cc < k;
    const char *Xl = reinterpret_cast<const char *>(X + kb + 2 * cc);
    double *Yl = Y + kb + 2 * cc;
    // accumulators: mine[i] = column h*HW + i (stored by this lane), other[i] = column (1-h)*HW + i (sent to the partner half)
    double mine[HW][2], other[HW][2];
#pragma unroll
    for (int i = 0; i < HW; i++) { mine[i][0] = mine[i][1] = other[i][0] = other[i][1] = 0.0; }
    const int vofs_mine = h * W + h * HW, vofs_other = h * W + (1 - h) * HW; // within a pair of rows of a val chunk

    // unit u = stripes [unit_lo(u), unit_lo(u+1)): at most 32 of them (the host picks nunits so)
    auto unit_lo = [&](const int u) { return u >= nunits ? L : min(L, (int)((double)u * ratio)); };
    // ---- chunk generator: every unit yields ceil(rows / 16) chunks (at least one), never straddling units
    int gu = wid, gP, gE, gnP = 0, gnE = 0;
    bool gfresh = true;
    gP = __ldg(&meta[unit_lo(gu)].pos); gE = __ldg(&meta[unit_lo(gu + 1)].pos);
    if (gu + nwarps < nunits) { gnP = __ldg(&meta[unit_lo(gu + nwarps)].pos); gnE = __ldg(&meta[unit_lo(gu + nwarps + 1)].pos); }
    int cu, cP, cn; // the chunk just generated: unit (-1: none left), first row, rows
    auto gen = [&]() {
        if (gP >= gE && !gfresh) {
            gu += nwarps; gP = gnP; gE = gnE; gfresh = true;
            if (gu + nwarps < nunits) { gnP = __ldg(&meta[unit_lo(gu + nwarps)].pos); gnE = __ldg(&meta[unit_lo(gu + nwarps + 1)].pos); }
        }
        if (gu < nunits) { cu = gu; cP = gP; cn = min(CR, gE - gP); gP += CR; gfresh = false; }
        else { cu = -1; cP = 0; cn = 0; }
    };
    auto load_idx = [&](const int P, const int n) { return lane < n ? __ldcs(desc + P + lane) : 0; };
    auto copy_val = [&](const int P, const int n, const int buf) { // rows [P, P+n) of val -> vs[buf], 16 bytes per lane and copy
        const double *src = val + (long long)P * W;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&vs[buf][0]);
#pragma unroll
        for (int i = lane; i < CR * W / 2; i += 32)
            if (i < n * HW) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (unsigned)i), "l"(src + 2 * i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto load_x = [&](const int idx, const int j) {
        const unsigned xi = (unsigned)__shfl_sync(0xffffffffu, idx, 2 * j + h);
        double2 r = make_double2(0.0, 0.0);
        if (colok) r = __ldg(reinterpret_cast<const double2 *>(Xl + (size_t)xi * ldx_bytes));
        return r;
    };

    // ---- prologue: chunks 0 and 1 described, their indices and values requested, chunk 0's X rows requested
    int u0c, P0c, u1c, P1c, n1c, u2c, P2c, n2c;
    gen(); u1c = cu; P1c = cP; n1c = cn;
    gen(); u2c = cu; P2c = cP; n2c = cn;
    int idx1 = load_idx(P1c, n1c), idx2 = load_idx(P2c, n2c);
    copy_val(P1c, n1c, 0);
    copy_val(P2c, n2c, 1);
    double2 xr[NP];
#pragma unroll
    for (int j = 0; j < NP; j++) xr[j] = load_x(idx1, j);

    // ---- consumer state.  segs: lane i holds the end (row) of stripe l0 + i of the current unit; nsegs etc.: the unit after it
    int cur_u = -1, l = 0, l0 = 0, ulast = 0, seg_end = DEAD, segs = 0;
    int nlo = unit_lo(wid), nlast = unit_lo(wid + 1);
    int nsegs = lane < nlast - nlo ? __ldg(&meta[nlo + 1 + lane].pos) : 0;
    int buf = 0;
    auto flush = [&]() { // stripe l is complete: Y[l*W + h*HW + i, panel] <- alpha * (mine + partner's other) + beta * Y
#pragma unroll
        for (int i = 0; i < HW; i++) {
            const double s0 = mine[i][0] + __shfl_xor_sync(0xffffffffu, other[i][0], 16);
            const double s1 = mine[i][1] + __shfl_xor_sync(0xffffffffu, other[i][1], 16);
            if (colok) {
                double2 *yp = reinterpret_cast<double2 *>(Yl + ((long long)l * W + h * HW + i) * ldy);
                double2 o = make_double2(alpha * s0, alpha * s1);
                if (beta != 0.0) { const double2 old = *yp; o.x += beta * old.x; o.y += beta * old.y; }
                *yp = o;
            }
            mine[i][0] = mine[i][1] = other[i][0] = other[i][1] = 0.0;
        }
    };

    for (;;) {
        // chunk queue advances: 0 = consumed now, 1 = X rows requested during this chunk, 2 = indices / values requested now
        u0c = u1c; P0c = P1c;
        u1c = u2c; P1c = P2c; n1c = n2c;
        gen(); u2c = cu; P2c = cP; n2c = cn;
        idx1 = idx2;
        idx2 = load_idx(P2c, n2c);
        __syncwarp(); // every lane is done with the buffer chunk 2's values go to (read one chunk ago)
        copy_val(P2c, n2c, buf == 0 ? 2 : buf - 1);
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncwarp();
        if (u0c != cur_u) { // a unit begins: store the stripes of the previous one that are still open, take over the new one's ends
            if (seg_end != DEAD) { while (l < ulast) { flush(); l++; } }
            if (u0c < 0) return;
            cur_u = u0c; l = l0 = nlo; ulast = nlast; segs = nsegs;
            nlo = unit_lo(cur_u + nwarps); nlast = unit_lo(cur_u + nwarps + 1);
            nsegs = lane < nlast - nlo ? __ldg(&meta[nlo + 1 + lane].pos) : 0;
            seg_end = __shfl_sync(0xffffffffu, segs, 0);
        }
        const double *vm = &vs[buf][vofs_mine], *vo = &vs[buf][vofs_other];
        auto fma_pair = [&](const double2 x, const int j) {
#pragma unroll
            for (int i = 0; i < HW; i += 2) {
                const double2 a = *reinterpret_cast<const double2 *>(vm + 2 * j * W + i), b = *reinterpret_cast<const double2 *>(vo + 2 * j * W + i);
                mine[i][0] = fma(a.x, x.x, mine[i][0]); mine[i][1] = fma(a.x, x.y, mine[i][1]);
                other[i][0] = fma(b.x, x.x, other[i][0]); other[i][1] = fma(b.x, x.y, other[i][1]);
                mine[i + 1][0] = fma(a.y, x.x, mine[i + 1][0]); mine[i + 1][1] = fma(a.y, x.y, mine[i + 1][1]);
                other[i + 1][0] = fma(b.y, x.x, other[i + 1][0]); other[i + 1][1] = fma(b.y, x.y, other[i + 1][1]);
            }
        };
#pragma unroll
        for (int j = 0; j < NP; j++) {
            const int pa = P0c + 2 * j;
            if (pa + 1 < seg_end) fma_pair(xr[j], j);
            else if (seg_end != DEAD) { // a stripe ends at row pa or pa + 1 (or the unit did): one row at a time
#pragma unroll 1
                for (int hh = 0; hh < 2; hh++) {
                    while (seg_end != DEAD && pa + hh >= seg_end) {
                        flush(); l++;
                        seg_end = l == ulast ? DEAD : __shfl_sync(0xffffffffu, segs, l - l0);
                    }
                    if (seg_end != DEAD && h == hh) fma_pair(xr[j], j);
                }
            }
            xr[j] = load_x(idx1, j); // pair j of the next chunk
        }
        buf = buf == 2 ? 0 : buf + 1;
    }
}

// ---- Adjoint SpMM fed by TMA row gathers (Float64, rows mode, uniform stripe width 8, panels of <= 32 right-hand sides) -------------
// Same row-stream decomposition as k_spmm_adj_stream, same FP64 tensor tiles as k_spmm_adj_dmma, but the gathered X rows never
// pass through the load/store unit as per-lane requests (the L1 data pipe is what bounds both of those kernels):
//   * one `cp.async.bulk.tensor.2d ... tile::gather4` fetches four arbitrary rows of X (row indices straight from desc) as a
//     [4][16] Float64 tile; the tensor map is encoded with a {16, 1} box and SWIZZLE_128B, so that an 8-row x 128-byte atom
//     holds rows r = 0..7 with their 16-byte chunks XORed by r;
//   * a k-step of the m8n8k4 tile takes rows {0, 3, 4, 7} or {1, 2, 5, 6} of the atom: the four lanes of a quad (contraction
//     slots t = 0..3) then read chunks whose XOR patterns differ in bits 1-2 -- every B-fragment LDS.64 is conflict-free
//     (2 wavefronts for 256 bytes; the same fragment costs 8 when loaded from global memory);
//   * the 8 x 8 values of the chunk's val rows arrive by one 1-D bulk copy on the same mbarrier (A fragments: 2-way conflicts).
// Every warp runs its own ring of S stages (8 rows each) with one mbarrier per stage: no CTA-wide synchronisation at all.
__device__ __forceinline__ void mbar_init(const unsigned bar, const unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(const unsigned bar, const unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(const unsigned bar, const unsigned parity)
{
    unsigned ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_gather4(const unsigned dst, const CUtensorMap *tm, const int c0, const int r0, const int r1, const int r2, const int r3, const unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(const unsigned dst, const void *src, const unsigned bytes, const unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

#ifndef VBC_TMA_S
#define VBC_TMA_S 4
#endif
#ifndef VBC_TMA_MINB
#define VBC_TMA_MINB 2
#endif
constexpr int TMA_S = VBC_TMA_S;                           // stages per warp
constexpr int TMA_CR = 8;                                  // rows per stage
constexpr int TMA_STAGE_X = 2048, TMA_STAGE_V = 512;       // bytes per stage: two swizzle atoms of X, 8 x 8 values
constexpr int TMA_WARP_BYTES = TMA_S * (TMA_STAGE_X + TMA_STAGE_V);
constexpr int TMA_SMEM_BYTES = 8 * TMA_WARP_BYTES + 8 * TMA_S * 8 + 8 * TMA_S * 16 + 1024; // + mbarriers + chunk queue + alignment slack

template <bool FULL>
__global__ void __launch_bounds__(256, VBC_TMA_MINB) k_spmm_adj_tma(const __grid_constant__ CUtensorMap tmX, const StripeMeta *__restrict__ meta,
                                                         const int *__restrict__ desc, const double *__restrict__ val, double *__restrict__ Y,
                                                         const long long ldy, const int L, const int nunits, const double ratio, const int k,
                                                         const int kb, const double alpha, const double beta)
{
    constexpr int S = TMA_S, CR = TMA_CR, W = 8;
    constexpr int DEAD = INT_MIN;
    extern __shared__ unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, wq = threadIdx.x >> 5;
    const int nwarps = (int)gridDim.x * 8, wid = (int)blockIdx.x * 8 + wq;
    const unsigned smem0 = ((unsigned)__cvta_generic_to_shared(smem_raw) + 1023u) & ~1023u;
    const unsigned xs = smem0 + (unsigned)wq * (S * TMA_STAGE_X);                        // [S][2 atoms][8 rows][128 B]
    const unsigned vsm = smem0 + 8u * S * TMA_STAGE_X + (unsigned)wq * (S * TMA_STAGE_V); // [S][8 rows][8 values]
    const unsigned bars = smem0 + 8u * S * (TMA_STAGE_X + TMA_STAGE_V) + (unsigned)wq * (S * 8);
    const unsigned queue = smem0 + 8u * S * (TMA_STAGE_X + TMA_STAGE_V) + 8u * S * 8 + (unsigned)wq * (S * 16); // chunk descriptors {unit, first row, rows, -}
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; s++) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (wid >= nunits) return;

    auto unit_lo = [&](const int u) { return u >= nunits ? L : min(L, (int)((double)u * ratio)); };
    // ---- chunk generator (as in k_spmm_adj_stream): every unit yields ceil(rows / 8) chunks, at least one
    int gu = wid, gP, gE, gnP = 0, gnE = 0;
    bool gfresh = true;
    gP = __ldg(&meta[unit_lo(gu)].pos); gE = __ldg(&meta[unit_lo(gu + 1)].pos);
    if (gu + nwarps < nunits) { gnP = __ldg(&meta[unit_lo(gu + nwarps)].pos); gnE = __ldg(&meta[unit_lo(gu + nwarps + 1)].pos); }
    int cu, cP, cn;
    auto gen = [&]() {
        if (gP >= gE && !gfresh) {
            gu += nwarps; gP = gnP; gE = gnE; gfresh = true;
            if (gu + nwarps < nunits) { gnP = __ldg(&meta[unit_lo(gu + nwarps)].pos); gnE = __ldg(&meta[unit_lo(gu + nwarps + 1)].pos); }
        }
        if (gu < nunits) { cu = gu; cP = gP; cn = min(CR, gE - gP); gP += CR; gfresh = false; }
        else { cu = -1; cP = 0; cn = 0; }
    };
    auto put_desc = [&](const int slot) { // chunk descriptor -> queue[slot] (every lane writes the same words)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(queue + 16u * (unsigned)slot), "r"(cu), "r"(cP), "r"(cn), "r"(0) : "memory");
    };
    auto get_desc = [&](const int slot, int &u, int &P, int &n) {
        int z;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u), "=r"(P), "=r"(n), "=r"(z) : "r"(queue + 16u * (unsigned)slot) : "memory");
    };
    auto load_idx = [&](const int P, const int n) { return lane < n ? __ldcs(desc + P + lane) : 0; };
    // requests of one chunk: lanes 0 and 4 each gather four rows (both column halves); lane 0 also copies the values and arms the barrier
    auto issue = [&](const int u, const int P, const int n, const int idx, const int stage) {
        const int i1 = __shfl_down_sync(0xffffffffu, idx, 1), i2 = __shfl_down_sync(0xffffffffu, idx, 2), i3 = __shfl_down_sync(0xffffffffu, idx, 3);
        if (u < 0) return;
        const unsigned bar = bars + 8u * (unsigned)stage;
        if (lane == 0) {
            mbar_expect_tx(bar, (unsigned)(TMA_STAGE_X + n * W * 8));
            if (n > 0) bulk_copy_g2s(vsm + (unsigned)stage * TMA_STAGE_V, val + (long long)P * W, (unsigned)(n * W * 8), bar);
        }
        if (lane == 0 || lane == 4) {
            const unsigned dst = xs + (unsigned)stage * TMA_STAGE_X + (lane ? 512u : 0u);
            tma_gather4(dst, &tmX, kb, idx, i1, i2, i3, bar);
            tma_gather4(dst + 1024u, &tmX, kb + 16, idx, i1, i2, i3, bar);
        }
    };

    // ---- prologue: chunks 0..S-1 described, 0..S-2 requested
    int idxP = 0, pu = -1, pP = 0, pn = 0; // chunk whose requests go out next (S-1 ahead of the one being multiplied)
#pragma unroll
    for (int q = 0; q < S; q++) {
        gen(); put_desc(q);
        const int idx = load_idx(cP, cn);
        if (q < S - 1) issue(cu, cP, cn, idx, q);
        else { idxP = idx; pu = cu; pP = cP; pn = cn; }
    }
    __syncwarp();

    // ---- consumer state (stripe ends of the unit in `segs`, one per lane)
    int cur_u = -1, l = 0, l0 = 0, ulast = 0, seg_end = DEAD, segs = 0;
    int nlo = unit_lo(wid), nlast = unit_lo(wid + 1);
    int nsegs = lane < nlast - nlo ? __ldg(&meta[nlo + 1 + lane].pos) : 0;
    double c[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) c[nt][0] = c[nt][1] = 0.0;
    auto flush = [&]() {
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            const int col = kb + nt * 8 + 2 * t;
            if (FULL || col < k) {
                double2 *yp = reinterpret_cast<double2 *>(Y + ((long long)l * W + g) * ldy + col);
                double2 o = make_double2(alpha * c[nt][0], alpha * c[nt][1]);
                if (beta != 0.0) { const double2 old = *yp; o.x += beta * old.x; o.y += beta * old.y; }
                *yp = o;
            }
            c[nt][0] = c[nt][1] = 0.0;
        }
    };
    // this lane's rows in the two k-steps of a chunk and the swizzled byte offsets of its B-fragment elements (n-tiles 0 / 1; 2 and 3: + 1024)
    const int r_ks0 = 2 * t + (t & 1), r_ks1 = 2 * t + 1 - (t & 1); // {0, 3, 4, 7} and {1, 2, 5, 6}
    const unsigned xo0 = (unsigned)(r_ks0 * 128 + ((((g >> 1) ^ r_ks0) & 7) << 4) + (g & 1) * 8);
    const unsigned xo1 = (unsigned)(r_ks1 * 128 + ((((g >> 1) ^ r_ks1) & 7) << 4) + (g & 1) * 8);
    const unsigned vo0 = (unsigned)(r_ks0 * 64 + g * 8), vo1 = (unsigned)(r_ks1 * 64 + g * 8);

    int stage = 0;
    unsigned parity = 0;
    for (;;) {
        int u0c, P0c, n0c;
        get_desc(stage, u0c, P0c, n0c);
        // the stage multiplied in the previous iteration is free: chunk (S-1 ahead)'s requests go there
        issue(pu, pP, pn, idxP, stage == 0 ? S - 1 : stage - 1);
        // describe the chunk S ahead and fetch its row indices (requested in the next iteration)
        gen(); put_desc(stage); // into the slot just read (issue() above synchronised the warp after the read)
        pu = cu; pP = cP; pn = cn;
        idxP = load_idx(cP, cn);
        if (u0c != cur_u) {
            if (seg_end != DEAD) { while (l < ulast) { flush(); l++; } }
            if (u0c < 0) return;
            cur_u = u0c; l = l0 = nlo; ulast = nlast; segs = nsegs;
            nlo = unit_lo(cur_u + nwarps); nlast = unit_lo(cur_u + nwarps + 1);
            nsegs = lane < nlast - nlo ? __ldg(&meta[nlo + 1 + lane].pos) : 0;
            seg_end = __shfl_sync(0xffffffffu, segs, 0);
        }
        const unsigned bar = bars + 8u * (unsigned)stage;
        for (unsigned spins = 0; !mbar_try_wait(bar, parity); spins++)
            if (spins > (1u << 24)) __trap(); // a request that never completes must not hang the GPU (try_wait itself blocks for a while)
        const unsigned xb = xs + (unsigned)stage * TMA_STAGE_X, vb = vsm + (unsigned)stage * TMA_STAGE_V;
        double a0, a1, b0[4], b1[4];
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a0) : "r"(vb + vo0) : "memory");
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a1) : "r"(vb + vo1) : "memory");
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            const unsigned ofs = (nt >> 1) * 1024u;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b0[nt]) : "r"(xb + ofs + ((nt & 1) ? (xo0 ^ 64u) : xo0)) : "memory");
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b1[nt]) : "r"(xb + ofs + ((nt & 1) ? (xo1 ^ 64u) : xo1)) : "memory");
        }
        if (seg_end > P0c + CR) { // all eight rows belong to the open stripe
#pragma unroll
            for (int nt = 0; nt < 4; nt++) dmma_m8n8k4(c[nt], a0, b0[nt]);
#pragma unroll
            for (int nt = 0; nt < 4; nt++) dmma_m8n8k4(c[nt], a1, b1[nt]);
        } else {
            int lo = P0c;
            while (seg_end != DEAD) {
                const int hi = min(seg_end, P0c + CR);
                if (hi > lo) { // rows [lo, hi) of the chunk belong to stripe l: everything else is masked on both operands
                    const bool m0 = P0c + r_ks0 >= lo && P0c + r_ks0 < hi, m1 = P0c + r_ks1 >= lo && P0c + r_ks1 < hi;
                    const double am0 = m0 ? a0 : 0.0, am1 = m1 ? a1 : 0.0;
#pragma unroll
                    for (int nt = 0; nt < 4; nt++) dmma_m8n8k4(c[nt], am0, m0 ? b0[nt] : 0.0);
#pragma unroll
                    for (int nt = 0; nt < 4; nt++) dmma_m8n8k4(c[nt], am1, m1 ? b1[nt] : 0.0);
                }
                if (seg_end > P0c + CR) break; // the stripe continues in the next chunk
                flush(); l++;
                seg_end = l == ulast ? DEAD : __shfl_sync(0xffffffffu, segs, l - l0);
                lo = hi;
            }
        }
        __syncwarp(); // every lane has read the stage before the next iteration overwrites it
        stage = stage == S - 1 ? 0 : stage + 1;
        if (stage == 0) parity ^= 1u;
    }
}


// ================================================= launch (host) ====================================================
            // one stripe width (4 or 8) over the whole matrix, rows mode, 16-byte aligned even-k panels: the row-stream kernel
            const int Wu = A->w_uniform;
            const bool streamable = MODE == DESC_ROWS && (Wu == 4 || Wu == 8) && A->nval == A->ndesc * Wu && A->n == (int64_t)L * Wu &&
                                    (k % 2) == 0 && (ldx % 2) == 0 && (ldy % 2) == 0 && ldx < (1ll << 28) && ((uintptr_t)X % 16) == 0 && ((uintptr_t)Y % 16) == 0;
            if ((A->opt_spmm_simt == 0 || A->opt_spmm_simt == 3) && streamable && Wu == 8) { // TMA-fed tensor tiles
                typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
                static EncodeFn encode = nullptr;
                static bool attr_set = false;
                if (!encode) {
                    cudaDriverEntryPointQueryResult qres;
                    void *fn = nullptr;
                    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) VBC_FAIL(VBC_ECUDA, "cuTensorMapEncodeTiled is not available");
                    encode = (EncodeFn)fn;
                }
                if (!attr_set) {
                    VBC_CUDA(cudaFuncSetAttribute(k_spmm_adj_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BYTES));
                    VBC_CUDA(cudaFuncSetAttribute(k_spmm_adj_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BYTES));
                    attr_set = true;
                }
                CUtensorMap tm;
                const cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)A->m}, gstr[1] = {(cuuint64_t)ldx * 8};
                const cuuint32_t box[2] = {16, 1}, estr[2] = {1, 1};
                const CUresult cr = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<Tv *>(X), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (cr != CUDA_SUCCESS) VBC_FAIL(VBC_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
                int64_t g2 = (int64_t)A->sm_count * VBC_TMA_MINB;
                const int64_t nw = g2 * 8;
                const double avg_rows = (double)A->ndesc / (double)L;
                int64_t U0 = (int64_t)(384.0 / (avg_rows > 1.0 ? avg_rows : 1.0) + 0.5);
                if (U0 < 1) U0 = 1;
                if (U0 > 24) U0 = 24;
                int64_t q = (L + nw * U0 / 2) / (nw * U0);
                if (q < 1) q = 1;
                while ((double)L / (double)(nw * q) > 30.0) q++;
                int64_t nunits = nw * q;
                if (nunits > L) { nunits = L; g2 = (nunits + 7) / 8; }
                const double ratio = (double)L / (double)nunits;
                for (int kb = 0; kb < k; kb += 32) {
                    if (k - kb >= 32) k_spmm_adj_tma<true><<<(unsigned)g2, 256, TMA_SMEM_BYTES, A->stream>>>(tm, A->d_meta, A->d_desc, (const double *)A->d_val, (double *)Y, ldy, L, (int)nunits, ratio, k, kb, (double)alpha, (double)beta);
                    else k_spmm_adj_tma<false><<<(unsigned)g2, 256, TMA_SMEM_BYTES, A->stream>>>(tm, A->d_meta, A->d_desc, (const double *)A->d_val, (double *)Y, ldy, L, (int)nunits, ratio, k, kb, (double)alpha, (double)beta);
                    A->launches++;
                }
                VBC_CUDA(cudaGetLastError());
                return VBC_OK;
            }
            if ((A->opt_spmm_simt == 0 || A->opt_spmm_simt == 4) && streamable) {
                int64_t g2 = (int64_t)A->sm_count * 2;
                const int64_t nw = g2 * 8;
                const double avg_rows = (double)A->ndesc / (double)L;
                int64_t U0 = (int64_t)(384.0 / (avg_rows > 1.0 ? avg_rows : 1.0) + 0.5);
                if (U0 < 1) U0 = 1;
                if (U0 > 24) U0 = 24; // a unit's stripe ends sit in one register per lane: at most 32 stripes
                int64_t q = (L + nw * U0 / 2) / (nw * U0);
                if (q < 1) q = 1;
                while ((double)L / (double)(nw * q) > 30.0) q++;
                int64_t nunits = nw * q;
                if (nunits > L) { nunits = L; g2 = (nunits + 7) / 8; }
                const double ratio = (double)L / (double)nunits;
                const unsigned ldxb = (unsigned)(ldx * 8);
                for (int kb = 0; kb < k; kb += 32) {
#define STREAM_LAUNCH(Wv, FULLv) k_spmm_adj_stream<Wv, FULLv><<<(unsigned)g2, 256, 0, A->stream>>>(A->d_meta, A->d_desc, (const double *)A->d_val, (const double *)X, ldxb, (double *)Y, ldy, L, (int)nunits, ratio, k, kb, (double)alpha, (double)beta)
                    const bool full = k - kb >= 32;
                    if (Wu == 8) { if (full) STREAM_LAUNCH(8, true); else STREAM_LAUNCH(8, false); }
                    else         { if (full) STREAM_LAUNCH(4, true); else STREAM_LAUNCH(4, false); }
#undef STREAM_LAUNCH
                    A->launches++;
                }
                VBC_CUDA(cudaGetLastError());
                return VBC_OK;
            }
#endif
