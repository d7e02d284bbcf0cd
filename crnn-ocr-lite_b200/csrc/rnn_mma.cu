// rnn_mma.cu -- cluster-resident Bidirectional GRU recurrence (forward + BPTT) with the recurrent matrix held in REGISTERS
// as tensor-core fragments (sm_100a).
//
// rnn_cluster.cu keeps the U shard of each CTA in shared memory and re-reads all 96 KB of it with LDS.128 every time step:
// ncu (profiles/r01, gru_fwd) shows the step bound by those reads (short-scoreboard 41 % of the stall samples, two warps per
// scheduler) -- 7 us per step although the step's math is only 0.2 MFLOP per CTA.  Here the same 8-CTA cluster / DSMEM exchange
// protocol is kept, but the per-step contraction runs on the tensor cores, transposed so that the 8 batch rows of the cluster are
// exactly the N=8 of mma.sync.m16n8k8 (tf32):
//     forward : out^T (96 gate columns x 8 rows) = Ushard^T (96 x 256) . h^T (256 x 8)
//     backward: dh^T  (256 units x 8 rows)       = Ushard   (256 x 96) . da^T (96 x 8)
// U never changes during the sequence, so every warp loads its A fragments ONCE: the tf32 "hi" part as 48 registers per thread and
// the residual "lo" part as 24 registers of packed bf16 -- 72 of the 128 registers of a 512-thread CTA, 144 KB of the SM's register
// file.  Per step a warp only reads the tiny h / da operand from shared memory (conflict-free B fragments) and issues 8-24 MMAs.
// fp32 fidelity: a.b = a_hi.b_hi + a_hi.b_lo + a_lo.b_hi (the dropped a_lo.b_lo term is 2^-22 relative), see mma2() below.
// tcgen05 is not used here on purpose: with M = 8 batch rows per step the problem is latency-, not throughput-bound, and the
// smem-descriptor / commit / tcgen05.ld round trip of one UMMA is longer than this whole register-resident mma.sync chain.
// Same math as rnn.cu / the oracle (Keras 2.2.2 GRUCell reset_after=False, hard_sigmoid).
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace {
constexpr int U = 256;        // hidden units
constexpr int NCTA = 8;       // CTAs per cluster
constexpr int UPC = U / NCTA; // units per CTA (32)
constexpr int RB = 8;         // batch rows per cluster (= N of the MMA)
constexpr int NT = 512;       // threads per CTA (16 warps)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
// 16-byte remote store that also signals the DESTINATION CTA's mbarrier (complete_tx of 16 bytes): data + flag in one operation
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, float4 v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(remote_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count)); }
__device__ __forceinline__ void bar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {   // bounded: a protocol bug traps instead of hanging the GPU
    uint32_t done = 0, polls = 0;
    long long t0 = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
        if (!done && (++polls & 255u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now; else if (now - t0 > 4000000000ll) __trap();
        }
    }
}
constexpr uint32_t XCHG_BYTES = NCTA * UPC * RB * 4;   // bytes every CTA receives per exchange (8 KB)

// Per-step global operands (input projections / saved gates) are prefetched PF steps ahead straight into a per-thread shared-memory
// ring with 4-byte cp.async: no register is tied up by a load in flight (with U occupying 72 registers, register-prefetched values were
// spilled by ptxas, which waits for the load -- ncu r1i: 39 % of all stall samples on that one long-scoreboard wait).
constexpr int RING = 4, PF = RING - 1;
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pf() { asm volatile("cp.async.wait_group %0;" ::"n"(PF) : "memory"); }

__device__ __forceinline__ uint32_t to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// fp32-faithful product on the legacy tensor path in TWO instructions per 16x8x8 block instead of the three of "3xTF32":
//   a.b = a_hi.b_hi  (tf32 m16n8k8)  +  [ a_hi.b_lo + a_lo.b_hi ]  (ONE bf16 m16n8k16)
// The two cross terms are 2^-11 smaller than the main term, so bf16 operands (8 bits) are enough for them (error <= ~2^-19 of the
// product); they share one k16 MMA by interleaving along k: slot 2t <- (a_hi', b_lo), slot 2t+1 <- (a_lo, b_hi').  The element
// (row, k) -> register mapping of the bf16 A/B fragments is then exactly that of the tf32 fragments, so a_hi' is just the upper half
// of the tf32 register (one PRMT) and the residuals are stored as packed bf16 pairs.  (HMMA.1688.F32.TF32 issues once per ~13 cycles
// per SM sub-partition on B200 -- ncu r1j: the tensor pipe was the floor of every step.)
struct AFrag { uint32_t hi[4]; uint32_t lo[2]; };   // hi: tf32 bit patterns; lo: bf16(a - hi) of elements (0,1) and (2,3)
__device__ __forceinline__ uint32_t pack_bf16(float lo_half, float hi_half) {   // result = {hi_half -> bits 31:16, lo_half -> bits 15:0}
    uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_half), "f"(lo_half)); return r;
}
__device__ __forceinline__ void afrag_set(AFrag& f, float a0, float a1, float a2, float a3) {
    const float a[4] = {a0, a1, a2, a3};
    float l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { f.hi[i] = to_tf32(a[i]); l[i] = a[i] - __uint_as_float(f.hi[i]); }
    f.lo[0] = pack_bf16(l[0], l[1]); f.lo[1] = pack_bf16(l[2], l[3]);
}
// B fragment (k8 x n8) of a [k][row] (row fastest, 8 floats per k) shared-memory operand: tf32 hi parts + packed {b_lo, b_hi'} pairs
struct BFrag { uint32_t h0, h1, x0, x1; };
__device__ __forceinline__ BFrag bfrag(const float* __restrict__ p, int k0, int gid, int tig) {
    const float b0 = p[(k0 + tig) * RB + gid], b1 = p[(k0 + tig + 4) * RB + gid];
    BFrag f;
    f.h0 = to_tf32(b0); f.h1 = to_tf32(b1);
    const float h0 = __uint_as_float(f.h0), h1 = __uint_as_float(f.h1);
    f.x0 = pack_bf16(b0 - h0, h0); f.x1 = pack_bf16(b1 - h1, h1);        // slot 2t: b_lo, slot 2t+1: b_hi'
    return f;
}
// acc_m += A_hi.B_hi ; acc_x += cross terms   (two accumulators: the dependent MMA chains stay short)
__device__ __forceinline__ void mma2(float (&am)[4], float (&ax)[4], const uint32_t (&hi)[4], uint32_t lo01, uint32_t lo23, const BFrag& b) {
    mma_tf32(am, hi, b.h0, b.h1);
    mma_bf16(ax, __byte_perm(hi[0], lo01, 0x5432), __byte_perm(hi[1], lo01, 0x7632), __byte_perm(hi[2], lo23, 0x5432), __byte_perm(hi[3], lo23, 0x7632), b.x0, b.x1);
}
__device__ __forceinline__ void mma2(float (&am)[4], float (&ax)[4], const AFrag& f, const BFrag& b) { mma2(am, ax, f.hi, f.lo[0], f.lo[1], b); }

// Broadcast one value per (unit, row) to the same offset of `buf` in all 8 CTAs, signalling each destination's barrier.  Rows are the
// fastest index (er = lane & 7): the lanes with er % 4 == 0 gather their 4-row group with shuffles -> one 16-byte st.async per destination.
__device__ __forceinline__ void push4_async(float* buf, uint64_t* bar, int off, float v, int er) {
    const float v1 = __shfl_down_sync(0xffffffffu, v, 1), v2 = __shfl_down_sync(0xffffffffu, v, 2), v3 = __shfl_down_sync(0xffffffffu, v, 3);
    if ((er & 3) == 0) {
        const float4 q = make_float4(v, v1, v2, v3);
        const uint32_t a = smem_addr(buf + off), b = smem_addr(bar);
#pragma unroll
        for (int c = 0; c < NCTA; ++c) st_async_v4(map_rank(a, c), q, map_rank(b, c));
    }
}

// ------------------------------------------------------------------------------------------------- forward
// smem: hT[2][256][8] | rhT[2][256][8] | partA[8][64][8] | partB[8][32][8] | 4 mbarriers
constexpr int FWD_SMEM_FLOATS = 2 * U * RB + 2 * U * RB + 8 * 64 * RB + 8 * 32 * RB + RING * 2 * NT + 16;

__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(NT, 1)
gru_fwd_mma_kernel(const float* __restrict__ xp, const float* __restrict__ U0, const float* __restrict__ U1,
                   float* __restrict__ hs, float* __restrict__ gates, int B, int T)
{ pdl_enter();
    extern __shared__ __align__(16) float sm[];
    float* hT = sm;                                 // [2][256][8]  h_{t-1} of all units (k-major, rows fastest)
    float* rhT = hT + 2 * U * RB;                   // [2][256][8]  r * h_{t-1}
    float* partA = rhT + 2 * U * RB;                // [8 k-slices][64 cols (z: 0..31, r: 32..63)][8 rows]
    float* partB = partA + 8 * 64 * RB;             // [8 k-slices][32 cols][8 rows]
    float* xring = partB + 8 * 32 * RB;             // [RING][2][NT] prefetched input projections of this thread's gate(s)
    uint64_t* barR = reinterpret_cast<uint64_t*>(xring + RING * 2 * NT);   // [2] r*h exchange of step s -> barR[s&1]
    uint64_t* barH = barR + 2;                                          // [2] h   exchange of step s -> barH[s&1]
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int dir = blockIdx.y, b0 = (blockIdx.x / NCTA) * RB;
    const float* Um = dir ? U1 : U0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
    // (tried: r gate alone before the exchange and the z gate's MMAs deferred into the r*h exchange window -- 8 fewer MMAs per warp on
    // the critical path, step time unchanged at 4.2 us: the step is a chain of many small latencies (two exchanges 27 %, MMA phases 34 %,
    // element-wise + push 20 %, barriers 7 %, ncu r1n), not MMA-bound.)
    // MMA identity: warp = (k-slice of 32, column group mp).  Phase A: the 32 columns of gate mp (0: z, 1: r) = 2 m-tiles;
    // phase B: columns [16 mp, 16 mp + 16) of the candidate gate = 1 m-tile.
    const int mp = warp & 1, ks = warp >> 1;

    AFrag fa[2][4], fb[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const float* r0 = Um + (size_t)(32 * ks + 8 * kk + tig) * (3 * U);
        const float* r1 = r0 + (size_t)4 * (3 * U);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int c = mp * U + crank * UPC + 16 * mt + gid;
            afrag_set(fa[mt][kk], __ldg(r0 + c), __ldg(r0 + c + 8), __ldg(r1 + c), __ldg(r1 + c + 8));
        }
        const int c = 2 * U + crank * UPC + 16 * mp + gid;
        afrag_set(fb[kk], __ldg(r0 + c), __ldg(r0 + c + 8), __ldg(r1 + c), __ldg(r1 + c + 8));
    }
    for (int i = tid; i < 2 * U * RB; i += NT) hT[i] = 0.f;
    if (tid == 0) {
        bar_init(&barR[0], 1); bar_init(&barR[1], 1); bar_init(&barH[0], 1); bar_init(&barH[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();                                  // every CTA's barriers and buffers exist before the first remote store

    // element identity: tid = gate (0: z, 1: r) x unit eu x row er (fastest); the z threads also own the candidate gate and h
    const int er = tid & 7, eu = (tid >> 3) & 31, isr = tid >> 8;
    const int j = crank * UPC + eu;              // global hidden unit
    const int b = b0 + er;
    const bool valid = b < B;
    const int bb = valid ? b : B - 1;

    // projections of the own gate (z or r) and, for the z threads, of the candidate gate: cp.async ring, PF steps ahead
    const float* xbase = xp + ((size_t)bb * T * 2 + dir) * (3 * U) + isr * U + j;
    auto prefetch = [&](int sp) {
        if (sp < T) {
            const float* x = xbase + (size_t)(dir ? T - 1 - sp : sp) * (2 * 3 * U);
            float* d = xring + (sp % RING) * (2 * NT) + tid;
            cp_async4(d, x);
            if (!isr) cp_async4(d + NT, x + 2 * U);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int sp = 0; sp < PF; ++sp) prefetch(sp);
    for (int s = 0; s < T; ++s) {
        const int t = dir ? T - 1 - s : s;
        const int cur = s & 1;
        const uint32_t ph = (s >> 1) & 1;
        const float* hcur = hT + cur * (U * RB);
        float* hnxt = hT + (cur ^ 1) * (U * RB);
        float* rhc = rhT + cur * (U * RB);
        prefetch(s + PF);
        if (s > 0) bar_wait(&barH[(s - 1) & 1], ((s - 1) >> 1) & 1);      // h_{t-1} of all 256 units has landed
        // ---- phase A: z / r pre-activation partials of the k-slice
        {
            float am[2][4] = {}, ax[2][4] = {};            // main and cross-term chains; 2 m-tiles interleave
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const BFrag bf = bfrag(hcur, 32 * ks + 8 * kk, gid, tig);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma2(am[mt], ax[mt], fa[mt][kk], bf);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                float* p = partA + ((ks * 64 + mp * 32 + 16 * mt + gid) * RB + 2 * tig);
                *reinterpret_cast<float2*>(p) = make_float2(am[mt][0] + ax[mt][0], am[mt][1] + ax[mt][1]);
                *reinterpret_cast<float2*>(p + 8 * RB) = make_float2(am[mt][2] + ax[mt][2], am[mt][3] + ax[mt][3]);
            }
        }
        const float hown = hcur[j * RB + er];
        cp_async_wait_pf();                                  // this thread's projections of step s are in its ring slot
        const float xg = xring[(s % RING) * (2 * NT) + tid];
        __syncthreads();
        float ag = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) ag += partA[q * 512 + tid];
        const float gv = hard_sigmoid(xg + ag);              // z (threads 0..255) or r (threads 256..511)
        if (tid == 0) bar_expect(&barR[cur], XCHG_BYTES);
        if (isr) {                                           // warps 8..15, warp-uniform
            push4_async(rhc, &barR[cur], j * RB + er, gv * hown, er);
            if (valid && gates) gates[(((size_t)b * T + t) * 2 + dir) * (3 * U) + U + j] = gv;
        }
        bar_wait(&barR[cur], ph);
        // ---- phase B: candidate pre-activation partials
        {
            float am[4] = {}, ax[4] = {};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma2(am, ax, fb[kk], bfrag(rhc, 32 * ks + 8 * kk, gid, tig));
            float* p = partB + ((ks * 32 + 16 * mp + gid) * RB + 2 * tig);
            *reinterpret_cast<float2*>(p) = make_float2(am[0] + ax[0], am[1] + ax[1]);
            *reinterpret_cast<float2*>(p + 8 * RB) = make_float2(am[2] + ax[2], am[3] + ax[3]);
        }
        __syncthreads();
        if (tid == 0) bar_expect(&barH[cur], XCHG_BYTES);
        if (!isr) {                                          // warps 0..7
            float ah = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) ah += partB[q * 256 + tid];
            const float hh = tanhf(xring[(s % RING) * (2 * NT) + NT + tid] + ah);
            const float hn = gv * hown + (1.f - gv) * hh;
            push4_async(hnxt, &barH[cur], j * RB + er, hn, er);
            if (valid) {
                const size_t o = ((size_t)b * T + t) * 2 + dir;
                hs[o * U + j] = hn;
                if (gates) { float* g = gates + o * (3 * U); g[j] = gv; g[2 * U + j] = hh; }
            }
        }
    }
    bar_wait(&barH[(T - 1) & 1], ((T - 1) >> 1) & 1);       // drain: all stores targeting this CTA have landed
    cluster.sync();                                         // nobody exits while a peer may still address its shared memory
}

// ------------------------------------------------------------------------------------------------- backward (BPTT)
// smem: da[3][32][8] | recvA[8][32][8] | recvB[8][32][8] | 2 mbarriers
constexpr int BWD_SMEM_FLOATS = 3 * UPC * RB + 2 * NCTA * UPC * RB + RING * 5 * (UPC * RB) + 8;

// Push this warp's 16(unit) x 8(row) partial tile into the owner CTA's receive buffer (slot of this source CTA): lane pairs swap
// halves so that every lane sends one 16-byte st.async (4 consecutive rows of one unit), each signalling 16 bytes on the owner's barrier.
__device__ __forceinline__ void push_tile(const float (&c)[4], float* recv_local, uint64_t* bar_local, int owner, int crank, int ubase, int gid, int tig) {
    const bool odd = tig & 1;
    const float x = __shfl_xor_sync(0xffffffffu, odd ? c[0] : c[2], 1);
    const float y = __shfl_xor_sync(0xffffffffu, odd ? c[1] : c[3], 1);
    const float4 q = odd ? make_float4(x, y, c[2], c[3]) : make_float4(c[0], c[1], x, y);
    const int u = ubase + gid + (odd ? 8 : 0);
    st_async_v4(map_rank(smem_addr(recv_local + crank * (UPC * RB) + u * RB + 4 * (tig >> 1)), owner), q, map_rank(smem_addr(bar_local), owner));
}

__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(NT, 1)
gru_bwd_mma_kernel(const float* __restrict__ dout, const float* __restrict__ hs, const float* __restrict__ gates,
                   const float* __restrict__ U0, const float* __restrict__ U1,
                   float* __restrict__ dxp, float* __restrict__ hprev_out, float* __restrict__ rh_out, int B, int T)
{ pdl_enter();
    extern __shared__ __align__(16) float sm[];
    float* da = sm;                                 // [3 gates z,r,h][32 units][8 rows]
    float* recvA = da + 3 * UPC * RB;               // [8 src][32][8]
    float* recvB = recvA + NCTA * UPC * RB;
    float* gring = recvB + NCTA * UPC * RB;         // [RING][5: z, r, hh, h_prev, dout][256] prefetched per-step operands of the element-wise threads
    uint64_t* barA = reinterpret_cast<uint64_t*>(gring + RING * 5 * (UPC * RB));   // d(r*h) reduce-scatter
    uint64_t* barB = barA + 1;                                                // dh_{t-1} reduce-scatter
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int dir = blockIdx.y, b0 = (blockIdx.x / NCTA) * RB;
    const float* Um = dir ? U1 : U0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;

    // warp w owns the 16 hidden units (rows of U) [16 w, 16 w + 16) for all 96 shard columns: 12 k-steps (0..7: gates z,r; 8..11: gate h)
    AFrag fu[12];
    {
        const float* r0 = Um + (size_t)(16 * warp + gid) * (3 * U);
        const float* r1 = r0 + (size_t)8 * (3 * U);
#pragma unroll
        for (int kk = 0; kk < 12; ++kk) {
            const int c = 8 * kk + tig;                                          // shard column (gate-major)
            const int cg0 = (c >> 5) * U + crank * UPC + (c & 31), cg1 = ((c + 4) >> 5) * U + crank * UPC + ((c + 4) & 31);
            afrag_set(fu[kk], __ldg(r0 + cg0), __ldg(r1 + cg0), __ldg(r0 + cg1), __ldg(r1 + cg1));
        }
    }
    if (tid == 0) {
        bar_init(barA, 1); bar_init(barB, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();

    const bool ew = tid < UPC * RB;                 // element-wise threads (warps 0..7): unit eu x row er
    const int er = tid & 7, eu = (tid >> 3) & 31;
    const int j = crank * UPC + eu;
    const int b = b0 + er;
    const bool valid = b < B;
    const int bb = valid ? b : B - 1;
    const int owner = warp >> 1, ubase = 16 * (warp & 1);

    float dh = 0.f;
    // the per-step operands (saved gates, h_{t-1}, upstream gradient) are prefetched PF steps ahead into a shared-memory ring (cp.async)
    constexpr int EW = UPC * RB;
    auto prefetch = [&](int s) {                    // s counts down; s < 0 -> empty group
        if (ew && s >= 0) {
            const int t = dir ? T - 1 - s : s;
            const int tp = dir ? t + 1 : t - 1;
            const size_t o = ((size_t)bb * T + t) * 2 + dir;
            const float* g = gates + o * (3 * U) + j;
            float* d = gring + (s % RING) * (5 * EW) + tid;
            cp_async4(d, g); cp_async4(d + EW, g + U); cp_async4(d + 2 * EW, g + 2 * U);
            if (s > 0) cp_async4(d + 3 * EW, hs + (((size_t)bb * T + tp) * 2 + dir) * U + j);
            cp_async4(d + 4 * EW, dout + o * U + j);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int q = 0; q < PF; ++q) prefetch(T - 1 - q);
    for (int s = T - 1; s >= 0; --s) {
        const int t = dir ? T - 1 - s : s;
        const uint32_t xph = (uint32_t)((T - 1 - s) & 1);       // phase parity of this step's two exchanges
        const size_t o = ((size_t)bb * T + t) * 2 + dir;
        prefetch(s - PF);
        cp_async_wait_pf();
        float z = 0.f, r = 0.f, hh = 0.f, hp = 0.f, dht = 0.f, dhn = 0.f, da_h = 0.f;
        if (ew) {
            const float* d = gring + (s % RING) * (5 * EW) + tid;
            z = d[0]; r = d[EW]; hh = d[2 * EW];
            hp = (s > 0) ? d[3 * EW] : 0.f;
            dht = (valid ? d[4 * EW] : 0.f) + dh;
            dhn = dht * z;
            da_h = dht * (1.f - z) * (1.f - hh * hh);
            da[(2 * UPC + eu) * RB + er] = da_h;
        }
        __syncthreads();
        // ---- d(r*h) partial of the warp's 16 units from the own 32 candidate-gate columns; reduce-scatter into recvA
        {
            float am[4] = {}, ax[4] = {};
#pragma unroll
            for (int kk = 8; kk < 12; ++kk) mma2(am, ax, fu[kk], bfrag(da, 8 * kk, gid, tig));
            const float c[4] = {am[0] + ax[0], am[1] + ax[1], am[2] + ax[2], am[3] + ax[3]};
            push_tile(c, recvA, barA, owner, crank, ubase, gid, tig);
        }
        if (tid == 0) bar_expect(barA, XCHG_BYTES);
        float da_z = 0.f, da_r = 0.f;
        if (ew) {
            bar_wait(barA, xph);
            float drh = 0.f;
#pragma unroll
            for (int c = 0; c < NCTA; ++c) drh += recvA[(c * UPC + eu) * RB + er];
            const float dz = dht * (hp - hh);
            const float dr = drh * hp;
            dhn = fmaf(drh, r, dhn);
            da_z = (z > 0.f && z < 1.f) ? 0.2f * dz : 0.f;
            da_r = (r > 0.f && r < 1.f) ? 0.2f * dr : 0.f;
            da[(0 * UPC + eu) * RB + er] = da_z;
            da[(1 * UPC + eu) * RB + er] = da_r;
        }
        __syncthreads();
        // ---- dh_{t-1} partial from the z and r gates (64 own columns); reduce-scatter into recvB
        {
            float am[4] = {}, ax[4] = {};
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) mma2(am, ax, fu[kk], bfrag(da, 8 * kk, gid, tig));
            const float c[4] = {am[0] + ax[0], am[1] + ax[1], am[2] + ax[2], am[3] + ax[3]};
            push_tile(c, recvB, barB, owner, crank, ubase, gid, tig);
        }
        if (tid == 0) bar_expect(barB, XCHG_BYTES);
        if (ew) {
            if (valid) {
                float* d = dxp + o * (3 * U);
                d[j] = da_z; d[U + j] = da_r; d[2 * U + j] = da_h;
                hprev_out[o * U + j] = hp;
                rh_out[o * U + j] = r * hp;
            }
            bar_wait(barB, xph);
#pragma unroll
            for (int c = 0; c < NCTA; ++c) dhn += recvB[(c * UPC + eu) * RB + er];
            dh = dhn;
        }
        // no trailing barrier: da[h] (written next) is not read by phase 2; recvA / recvB are only rewritten by peers after they
        // have consumed this step's pushes, which this CTA issues after the reads above (see the protocol notes in rnn_cluster.cu)
    }
    cluster.sync();               // nobody exits while a peer may still address its shared memory
}

// ================================================================================================= LSTM
// Keras 2.2.2 LSTMCell: kernel columns [i, f, c, o], i/f/o = hard_sigmoid, c' = f*c + i*tanh(.), h' = o*tanh(c').
// One exchange per step (no r*h dependency): every CTA owns 32 units x 4 gates = 128 columns of U (256 x 1024).  The tf32 hi parts of
// its 16 A fragments take 64 registers per thread; the bf16 residual pairs do not fit beside them and live in shared memory in
// fragment order (one conflict-free LDS.64 per fragment and step).
constexpr int LG = 4;                 // gates
constexpr int LC = LG * UPC;          // gate columns per CTA (128)
constexpr int LFR = 16;               // A fragments per warp

// ---- forward.  warp = (k-quarter kq: 64 k = 8 k-steps, gate g: 2 m-tiles of 16 units)
// smem: hT[2][256][8] | part[4][128][8] | lo[16][512] uint2 | xring[RING][4][256] | 2 mbarriers
constexpr int LHS = 4;                // of the 16 fragments, this many keep their hi parts in shared memory too (12 x 4 = 48 registers stay resident:
                                      // with all 64 in registers ptxas spilled ~30 of them to local memory and reloaded them every step)
constexpr int LF_SMEM_BYTES = 4 * (2 * U * RB + 4 * LC * RB + RING * 4 * (UPC * RB)) + 8 * LFR * NT + 16 * LHS * NT + 32;

__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(NT, 1)
lstm_fwd_mma_kernel(const float* __restrict__ xp, const float* __restrict__ U0, const float* __restrict__ U1,
                    float* __restrict__ hs, float* __restrict__ gates, int B, int T)
{ pdl_enter();
    extern __shared__ __align__(16) float sm[];
    float* hT = sm;                                 // [2][256][8]
    float* part = hT + 2 * U * RB;                  // [4 k-quarters][128 cols (gate-major)][8 rows]
    uint4* his = reinterpret_cast<uint4*>(part + 4 * LC * RB);     // [4 frags][512 threads]  hi parts of k-steps 6, 7
    uint2* los = reinterpret_cast<uint2*>(his + LHS * NT);         // [16 frags][512 threads]
    float* xring = reinterpret_cast<float*>(los + LFR * NT);       // [RING][4 gates][256]
    uint64_t* barH = reinterpret_cast<uint64_t*>(xring + RING * 4 * (UPC * RB));   // [2]
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int dir = blockIdx.y, b0 = (blockIdx.x / NCTA) * RB;
    const float* Um = dir ? U1 : U0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
    const int kq = warp & 3, g = warp >> 2;

    uint32_t fh[2][6][4];                           // tf32 hi parts of the A fragments (m-tile, k-step 0..5); k-steps 6, 7 live in `his`
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        const float* r0 = Um + (size_t)(64 * kq + 8 * kk + tig) * (LG * U);
        const float* r1 = r0 + (size_t)4 * (LG * U);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int c = g * U + crank * UPC + 16 * mt + gid;
            AFrag f; afrag_set(f, __ldg(r0 + c), __ldg(r0 + c + 8), __ldg(r1 + c), __ldg(r1 + c + 8));
            if (kk < 6) {
#pragma unroll
                for (int i = 0; i < 4; ++i) fh[mt][kk][i] = f.hi[i];
            } else his[(mt * 2 + kk - 6) * NT + tid] = make_uint4(f.hi[0], f.hi[1], f.hi[2], f.hi[3]);
            los[(mt * 8 + kk) * NT + tid] = make_uint2(f.lo[0], f.lo[1]);
        }
    }
    for (int i = tid; i < 2 * U * RB; i += NT) hT[i] = 0.f;
    if (tid == 0) {
        bar_init(&barH[0], 1); bar_init(&barH[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();

    const bool ew = tid < UPC * RB;                 // element-wise threads: unit eu x row er
    const int er = tid & 7, eu = (tid >> 3) & 31;
    const int j = crank * UPC + eu;
    const int b = b0 + er;
    const bool valid = b < B;
    const int bb = valid ? b : B - 1;
    constexpr int EW = UPC * RB;
    const float* xbase = xp + ((size_t)bb * T * 2 + dir) * (LG * U) + j;
    auto prefetch = [&](int sp) {
        if (ew && sp < T) {
            const float* x = xbase + (size_t)(dir ? T - 1 - sp : sp) * (2 * LG * U);
            float* d = xring + (sp % RING) * (4 * EW) + tid;
#pragma unroll
            for (int q = 0; q < LG; ++q) cp_async4(d + q * EW, x + q * U);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int sp = 0; sp < PF; ++sp) prefetch(sp);
    float cst = 0.f;                                // cell state of (unit eu, row er)
    for (int s = 0; s < T; ++s) {
        const int t = dir ? T - 1 - s : s;
        const int cur = s & 1;
        const float* hcur = hT + cur * (U * RB);
        float* hnxt = hT + (cur ^ 1) * (U * RB);
        prefetch(s + PF);
        if (s > 0) bar_wait(&barH[(s - 1) & 1], ((s - 1) >> 1) & 1);      // h_{t-1} of all 256 units has landed
        {
            float am[2][4] = {}, ax[2][4] = {};
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const BFrag bf = bfrag(hcur, 64 * kq + 8 * kk, gid, tig);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const uint2 lo = los[(mt * 8 + kk) * NT + tid];
                    if (kk < 6) mma2(am[mt], ax[mt], fh[mt][kk], lo.x, lo.y, bf);
                    else {
                        const uint4 hv = his[(mt * 2 + kk - 6) * NT + tid];
                        const uint32_t hi[4] = {hv.x, hv.y, hv.z, hv.w};
                        mma2(am[mt], ax[mt], hi, lo.x, lo.y, bf);
                    }
                }
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                float* p = part + ((kq * LC + g * UPC + 16 * mt + gid) * RB + 2 * tig);
                *reinterpret_cast<float2*>(p) = make_float2(am[mt][0] + ax[mt][0], am[mt][1] + ax[mt][1]);
                *reinterpret_cast<float2*>(p + 8 * RB) = make_float2(am[mt][2] + ax[mt][2], am[mt][3] + ax[mt][3]);
            }
        }
        cp_async_wait_pf();
        __syncthreads();
        if (tid == 0) bar_expect(&barH[cur], XCHG_BYTES);
        if (ew) {                                   // warps 0..7, warp-uniform
            float a[LG];
#pragma unroll
            for (int q = 0; q < LG; ++q) {
                float v = xring[(s % RING) * (4 * EW) + q * EW + tid];
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) v += part[k4 * (LC * RB) + q * EW + tid];
                a[q] = v;
            }
            const float ig = hard_sigmoid(a[0]), fg = hard_sigmoid(a[1]), gg = tanhf(a[2]), og = hard_sigmoid(a[3]);
            cst = fg * cst + ig * gg;
            const float hn = og * tanhf(cst);
            push4_async(hnxt, &barH[cur], j * RB + er, hn, er);
            if (valid) {
                const size_t o = ((size_t)b * T + t) * 2 + dir;
                hs[o * U + j] = hn;
                if (gates) { float* gp = gates + o * (5 * U); gp[j] = ig; gp[U + j] = fg; gp[2 * U + j] = gg; gp[3 * U + j] = og; gp[4 * U + j] = cst; }
            }
        }
        // no second barrier: `part` is rewritten only after the next barH wait, which completes only after every element-wise
        // thread of this CTA has pushed (i.e. has finished reading `part`)
    }
    bar_wait(&barH[(T - 1) & 1], ((T - 1) >> 1) & 1);
    cluster.sync();
}

// ---- backward.  warp w owns hidden units [16 w, 16 w + 16) for all 128 shard columns (16 k-steps, gate-major)
// smem: da[2][4][32][8] | recv[2][8][32][8] | lo[16][512] uint2 | gring[RING][8][256] | 2 mbarriers
constexpr int LB_SMEM_BYTES = 4 * (2 * LC * RB + 2 * NCTA * UPC * RB + RING * 8 * (UPC * RB)) + 8 * LFR * NT + 16 * LHS * NT + 32;

__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(NT, 1)
lstm_bwd_mma_kernel(const float* __restrict__ dout, const float* __restrict__ hs, const float* __restrict__ gates,
                    const float* __restrict__ U0, const float* __restrict__ U1,
                    float* __restrict__ dxp, float* __restrict__ hprev_out, int B, int T)
{ pdl_enter();
    extern __shared__ __align__(16) float sm[];
    float* da = sm;                                 // [2][4 gates i,f,g,o][32 units][8 rows]
    float* recv = da + 2 * LC * RB;                 // [2][8 src][32][8]
    uint4* his = reinterpret_cast<uint4*>(recv + 2 * NCTA * UPC * RB);   // hi parts of k-steps 12..15
    uint2* los = reinterpret_cast<uint2*>(his + LHS * NT);
    float* gring = reinterpret_cast<float*>(los + LFR * NT);       // [RING][8: i, f, g, o, c, c_prev, h_prev, dout][256]
    uint64_t* bar = reinterpret_cast<uint64_t*>(gring + RING * 8 * (UPC * RB));   // [2]
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int dir = blockIdx.y, b0 = (blockIdx.x / NCTA) * RB;
    const float* Um = dir ? U1 : U0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;

    uint32_t fh[LFR - LHS][4];
    {
        const float* r0 = Um + (size_t)(16 * warp + gid) * (LG * U);
        const float* r1 = r0 + (size_t)8 * (LG * U);
#pragma unroll
        for (int kk = 0; kk < LFR; ++kk) {
            const int c = 8 * kk + tig;
            const int cg0 = (c >> 5) * U + crank * UPC + (c & 31), cg1 = ((c + 4) >> 5) * U + crank * UPC + ((c + 4) & 31);
            AFrag f; afrag_set(f, __ldg(r0 + cg0), __ldg(r1 + cg0), __ldg(r0 + cg1), __ldg(r1 + cg1));
            if (kk < LFR - LHS) {
#pragma unroll
                for (int i = 0; i < 4; ++i) fh[kk][i] = f.hi[i];
            } else his[(kk - (LFR - LHS)) * NT + tid] = make_uint4(f.hi[0], f.hi[1], f.hi[2], f.hi[3]);
            los[kk * NT + tid] = make_uint2(f.lo[0], f.lo[1]);
        }
    }
    if (tid == 0) {
        bar_init(&bar[0], 1); bar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();

    const bool ew = tid < UPC * RB;
    const int er = tid & 7, eu = (tid >> 3) & 31;
    const int j = crank * UPC + eu;
    const int b = b0 + er;
    const bool valid = b < B;
    const int bb = valid ? b : B - 1;
    const int owner = warp >> 1, ubase = 16 * (warp & 1);
    constexpr int EW = UPC * RB;
    auto prefetch = [&](int s) {                    // s counts down; s < 0 -> empty group
        if (ew && s >= 0) {
            const int t = dir ? T - 1 - s : s;
            const int tp = dir ? t + 1 : t - 1;
            const size_t o = ((size_t)bb * T + t) * 2 + dir;
            const float* gp = gates + o * (5 * U) + j;
            float* d = gring + (s % RING) * (8 * EW) + tid;
#pragma unroll
            for (int q = 0; q < 5; ++q) cp_async4(d + q * EW, gp + q * U);
            if (s > 0) {
                const size_t op = ((size_t)bb * T + tp) * 2 + dir;
                cp_async4(d + 5 * EW, gates + op * (5 * U) + 4 * U + j);
                cp_async4(d + 6 * EW, hs + op * U + j);
            }
            cp_async4(d + 7 * EW, dout + o * U + j);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int q = 0; q < PF; ++q) prefetch(T - 1 - q);
    float dh = 0.f, dc = 0.f;
    for (int s = T - 1; s >= 0; --s) {
        const int n = T - 1 - s;                    // step counter
        const int buf = n & 1;
        const uint32_t ph = (uint32_t)((n >> 1) & 1);
        const int t = dir ? T - 1 - s : s;
        float* dab = da + buf * (LC * RB);
        float* rcv = recv + buf * (NCTA * UPC * RB);
        prefetch(s - PF);
        cp_async_wait_pf();
        float sa[LG] = {0.f, 0.f, 0.f, 0.f}, hp = 0.f;
        if (ew) {
            const float* d = gring + (s % RING) * (8 * EW) + tid;
            const float ig = d[0], fg = d[EW], gg = d[2 * EW], og = d[3 * EW], cc = d[4 * EW];
            const float cp = (s > 0) ? d[5 * EW] : 0.f;
            hp = (s > 0) ? d[6 * EW] : 0.f;
            const float dht = (valid ? d[7 * EW] : 0.f) + dh;
            const float tc = tanhf(cc);
            const float dog = dht * tc;
            const float dct = dc + dht * og * (1.f - tc * tc);
            sa[0] = (ig > 0.f && ig < 1.f) ? 0.2f * dct * gg : 0.f;
            sa[1] = (fg > 0.f && fg < 1.f) ? 0.2f * dct * cp : 0.f;
            sa[2] = dct * ig * (1.f - gg * gg);
            sa[3] = (og > 0.f && og < 1.f) ? 0.2f * dog : 0.f;
            dc = dct * fg;
#pragma unroll
            for (int q = 0; q < LG; ++q) dab[(q * UPC + eu) * RB + er] = sa[q];
        }
        __syncthreads();
        {
            float am[4] = {}, ax[4] = {};
#pragma unroll
            for (int kk = 0; kk < LFR; ++kk) {
                const uint2 lo = los[kk * NT + tid];
                if (kk < LFR - LHS) mma2(am, ax, fh[kk], lo.x, lo.y, bfrag(dab, 8 * kk, gid, tig));
                else {
                    const uint4 hv = his[(kk - (LFR - LHS)) * NT + tid];
                    const uint32_t hi[4] = {hv.x, hv.y, hv.z, hv.w};
                    mma2(am, ax, hi, lo.x, lo.y, bfrag(dab, 8 * kk, gid, tig));
                }
            }
            const float c[4] = {am[0] + ax[0], am[1] + ax[1], am[2] + ax[2], am[3] + ax[3]};
            push_tile(c, rcv, &bar[buf], owner, crank, ubase, gid, tig);
        }
        if (tid == 0) bar_expect(&bar[buf], XCHG_BYTES);
        if (ew) {
            if (valid) {
                const size_t o = ((size_t)b * T + t) * 2 + dir;
                float* d = dxp + o * (LG * U);
#pragma unroll
                for (int q = 0; q < LG; ++q) d[q * U + j] = sa[q];
                hprev_out[o * U + j] = hp;
            }
            bar_wait(&bar[buf], ph);
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < NCTA; ++c) acc += rcv[(c * UPC + eu) * RB + er];
            dh = acc;
        }
        // da / recv are double-buffered: the buffers of this step are rewritten two steps later, after the barrier of the next step
    }
    cluster.sync();
}
}  // namespace

int launch_gru_fwd_mma(const float* xp, const float* U0, const float* U1, float* hs, float* gates, int B, int T, cudaStream_t st)
{
    g_crnn_family = CRNN_FAM_RNN_MMA;
    if (B <= 0) return CRNN_OK;
    const size_t smem = sizeof(float) * FWD_SMEM_FLOATS;
    static bool configured = false;
    if (!configured) { CUDA_TRY(cudaFuncSetAttribute(gru_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = true; }
    dim3 grid(NCTA * ceil_div(B, RB), 2);
    (void)crnn_launch(gru_fwd_mma_kernel, grid, NT, smem, st, xp, U0, U1, hs, gates, B, T);
    crnn_pdl_mark_sparse(st);
    LAUNCH_CHECK();
    return CRNN_OK;
}

int launch_gru_bwd_mma(const float* dout, const float* hs, const float* gates, const float* U0, const float* U1,
                       float* dxp, float* hprev, float* rh, int B, int T, cudaStream_t st)
{
    g_crnn_family = CRNN_FAM_RNN_MMA;
    if (B <= 0) return CRNN_OK;
    const size_t smem = sizeof(float) * BWD_SMEM_FLOATS;
    static bool configured = false;
    if (!configured) { CUDA_TRY(cudaFuncSetAttribute(gru_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = true; }
    dim3 grid(NCTA * ceil_div(B, RB), 2);
    (void)crnn_launch(gru_bwd_mma_kernel, grid, NT, smem, st, dout, hs, gates, U0, U1, dxp, hprev, rh, B, T);
    crnn_pdl_mark_sparse(st);
    LAUNCH_CHECK();
    return CRNN_OK;
}

int launch_lstm_fwd_mma(const float* xp, const float* U0, const float* U1, float* hs, float* gates, int B, int T, cudaStream_t st)
{
    g_crnn_family = CRNN_FAM_RNN_MMA;
    if (B <= 0) return CRNN_OK;
    static bool configured = false;
    if (!configured) { CUDA_TRY(cudaFuncSetAttribute(lstm_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LF_SMEM_BYTES)); configured = true; }
    dim3 grid(NCTA * ceil_div(B, RB), 2);
    (void)crnn_launch(lstm_fwd_mma_kernel, grid, NT, LF_SMEM_BYTES, st, xp, U0, U1, hs, gates, B, T);
    crnn_pdl_mark_sparse(st);
    LAUNCH_CHECK();
    return CRNN_OK;
}

int launch_lstm_bwd_mma(const float* dout, const float* hs, const float* gates, const float* U0, const float* U1,
                        float* dxp, float* hprev, int B, int T, cudaStream_t st)
{
    g_crnn_family = CRNN_FAM_RNN_MMA;
    if (B <= 0) return CRNN_OK;
    static bool configured = false;
    if (!configured) { CUDA_TRY(cudaFuncSetAttribute(lstm_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LB_SMEM_BYTES)); configured = true; }
    dim3 grid(NCTA * ceil_div(B, RB), 2);
    (void)crnn_launch(lstm_bwd_mma_kernel, grid, NT, LB_SMEM_BYTES, st, dout, hs, gates, U0, U1, dxp, hprev, B, T);
    crnn_pdl_mark_sparse(st);
    LAUNCH_CHECK();
    return CRNN_OK;
}
