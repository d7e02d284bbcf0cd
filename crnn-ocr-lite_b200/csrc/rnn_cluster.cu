// rnn_cluster.cu -- cluster-resident Bidirectional GRU recurrence (forward + BPTT) for sm_100a.
//
// The recurrent matrix U (256 x 768 fp32 = 768 KB per direction) does not fit one SM, so v1 (rnn.cu) streamed it from
// L2 every step.  Here one thread-block CLUSTER of 8 CTAs owns U for one direction: CTA c keeps the columns of its 32
// hidden units for all three gates (256 x 96 floats = 96 KB) resident in shared memory for the whole sequence, handles 8
// batch rows, and the CTAs exchange only the tiny per-step state through distributed shared memory:
//   forward : every CTA needs the full h_{t-1} (and r*h for the candidate gate) of its 8 rows -> each CTA pushes the 32x8
//             values it produced into all 8 CTAs' buffers (st.shared::cluster), one cluster barrier per exchange
//             (2 per step: after r*h, after h_t).
//   backward: dh_{t-1} = da @ U^T contracts over gate columns, i.e. every CTA produces a PARTIAL result for all 256 units
//             from its own columns -> reduce-scatter through DSMEM (each CTA pushes the 32-unit slice owned by CTA c' into
//             c's receive buffer), 2 per step.
// 16 clusters (8 batch-slices x 2 directions) = 128 SMs run concurrently at B=64.  Same math as rnn.cu / the oracle
// (Keras 2.2.2 GRUCell reset_after=False, hard_sigmoid); only the fp32 summation order differs.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace {
constexpr int U = 256;        // hidden units
constexpr int NCTA = 8;       // CTAs per cluster
constexpr int UPC = U / NCTA; // units per CTA (32)
constexpr int RB = 8;         // batch rows per cluster
constexpr int GC = 3 * UPC;   // gate columns per CTA (96)

__device__ __forceinline__ void fma44(float (&acc)[4][4], const float4 a, const float4 b) {
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
// 16-byte remote store that also signals the DESTINATION CTA's mbarrier (complete_tx of 16 bytes): data + flag in one operation,
// so the per-step exchange needs no cluster-wide barrier / release fence (ncu: ERRBAR + UCGABAR were ~25 % of the stall samples).
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
        if (!done && (++polls & 255u) == 0) {            // the clock is only consulted on the (never taken in a healthy run) slow path
            const long long now = clock64();
            if (t0 == 0) t0 = now; else if (now - t0 > 4000000000ll) __trap();
        }
    }
}
constexpr uint32_t XCHG_BYTES = NCTA * UPC * RB * 4;   // bytes every CTA receives per exchange (8 KB)

// Broadcast one value per (unit, row) to the same offset of `buf` in all 8 CTAs, signalling each destination's barrier.
__device__ __forceinline__ void push4_async(float* buf, uint64_t* bar, int off, float v, int er) {
    const float v1 = __shfl_down_sync(0xffffffffu, v, 1), v2 = __shfl_down_sync(0xffffffffu, v, 2), v3 = __shfl_down_sync(0xffffffffu, v, 3);
    if ((er & 3) == 0) {
        const float4 q = make_float4(v, v1, v2, v3);
        const uint32_t a = smem_addr(buf + off), b = smem_addr(bar);
#pragma unroll
        for (int c = 0; c < NCTA; ++c) st_async_v4(map_rank(a, c), q, map_rank(b, c));
    }
}

// Broadcast one value per (unit, row) to the same offset of `buf` in all 8 CTAs.  Rows are the fastest index (er = lane & 7), so the
// lanes with er % 4 == 0 gather their 4-row group with shuffles and issue ONE 16-byte st.shared::cluster per destination
// (512 instead of 2048 remote stores per exchange and CTA).
__device__ __forceinline__ void push4(cg::cluster_group& cluster, float* buf, int off, float v, int er) {
    const float v1 = __shfl_down_sync(0xffffffffu, v, 1), v2 = __shfl_down_sync(0xffffffffu, v, 2), v3 = __shfl_down_sync(0xffffffffu, v, 3);
    if ((er & 3) == 0) {
        const float4 q = make_float4(v, v1, v2, v3);
#pragma unroll
        for (int c = 0; c < NCTA; ++c) *reinterpret_cast<float4*>(cluster.map_shared_rank(buf, c) + off) = q;
    }
}

// ------------------------------------------------------------------------------------------------- forward
// smem: Us[256][96] | hT[2][256][8] | rhT[256][8] | part[4096]
constexpr int FWD_SMEM_FLOATS = U * GC + 2 * U * RB + 2 * U * RB + 2 * 4096 + 16;   // + 4 mbarriers

__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(256, 1)
gru_fwd_cluster_kernel(const float* __restrict__ xp, const float* __restrict__ U0, const float* __restrict__ U1,
                       float* __restrict__ hs, float* __restrict__ gates, int B, int T)
{
    extern __shared__ __align__(16) float sm[];
    float* Us = sm;
    float* hT = Us + U * GC;
    float* rhT = hT + 2 * U * RB;                  // [2][256][8]
    float* part = rhT + 2 * U * RB;                // phase A partials
    float* partB = part + 4096;                    // phase B partials (own buffer: two __syncthreads per step instead of four)
    uint64_t* barR = reinterpret_cast<uint64_t*>(partB + 4096);   // [2] r*h exchange of step s -> barR[s&1]
    uint64_t* barH = barR + 2;                                    // [2] h   exchange of step s -> barH[s&1]
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int dir = blockIdx.y, b0 = (blockIdx.x / NCTA) * RB;
    const float* Um = dir ? U1 : U0;
    const int tid = threadIdx.x;

    // U shard -> shared memory: 6144 float4, 24 per thread, 8 independent loads in flight (the scalar loop this replaces waited a full
    // memory round trip per element: 43 us = 9 % of the launch, ncu r1f)
#pragma unroll
    for (int it = 0; it < 24; it += 8) {
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int idx = tid + 256 * (it + q), k = idx / 24, r = idx - k * 24, g = r >> 3, u4 = r & 7;
            v[q] = __ldg(reinterpret_cast<const float4*>(Um + (size_t)k * (3 * U) + g * U + crank * UPC + u4 * 4));
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int idx = tid + 256 * (it + q), k = idx / 24, r = idx - k * 24;
            *reinterpret_cast<float4*>(Us + k * GC + r * 4) = v[q];
        }
    }
    for (int i = tid; i < 2 * U * RB; i += 256) hT[i] = 0.f;
    if (tid == 0) {
        bar_init(&barR[0], 1); bar_init(&barR[1], 1); bar_init(&barH[0], 1); bar_init(&barH[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();                                  // every CTA's barriers and buffers exist before the first remote store

    // epilogue identity: row er (fastest) x unit eu  -> consecutive threads write consecutive floats of hT[unit][row]
    const int er = tid & 7, eu = tid >> 3;
    const int j = crank * UPC + eu;              // global hidden unit
    const int b = b0 + er;
    const bool valid = b < B;
    const int bb = valid ? b : B - 1;
    // compute identities
    const int lane = tid & 31, warp = tid >> 5;
    const int a_rt = lane >> 4, a_ct = lane & 15;            // phase A: warp = k-slice (32 k), 2 row-tiles x 16 col-tiles
    const int b_ks = tid >> 4, b_rt = (tid >> 3) & 1, b_ct = tid & 7;   // phase B: 16 k-slices (16 k), 2 x 8 tiles

    float xz, xr, xh;
    {
        const int t0 = dir ? T - 1 : 0;
        const float* x = xp + (((size_t)bb * T + t0) * 2 + dir) * (3 * U);
        xz = x[j]; xr = x[U + j]; xh = x[2 * U + j];
    }
    // Per step: h_{t-1} lives in hT[s&1]; r*h goes through rhT[s&1]; h_t is written into hT[(s+1)&1] of every CTA.
    // Exchanges are st.async (data + complete_tx on the destination's mbarrier); the double buffers are protected by the data
    // dependencies themselves (a CTA can only run one exchange ahead of the slowest CTA of its cluster).
    for (int s = 0; s < T; ++s) {
        const int t = dir ? T - 1 - s : s;
        const int cur = s & 1;
        const uint32_t ph = (s >> 1) & 1;
        const float* hcur = hT + cur * (U * RB);
        float* hnxt = hT + (cur ^ 1) * (U * RB);
        float* rhc = rhT + cur * (U * RB);
        float nxz = 0.f, nxr = 0.f, nxh = 0.f;             // next step's projections: issued now, consumed one step later
        if (s + 1 < T) {
            const int tn = dir ? T - 2 - s : s + 1;
            const float* x = xp + (((size_t)bb * T + tn) * 2 + dir) * (3 * U);
            nxz = __ldg(x + j); nxr = __ldg(x + U + j); nxh = __ldg(x + 2 * U + j);
        }
        if (s > 0) bar_wait(&barH[(s - 1) & 1], ((s - 1) >> 1) & 1);      // h_{t-1} of all 256 units has landed
        // ---- phase A: z,r pre-activations of the own 32 units: part[ks][col 0..63][row 0..7]
        {
            float acc[4][4] = {};
            const float* hp = hcur + (warp * 32) * RB + a_rt * 4;
            const float* up = Us + (warp * 32) * GC + a_ct * 4;
#pragma unroll 8
            for (int kk = 0; kk < 32; ++kk)
                fma44(acc, *reinterpret_cast<const float4*>(hp + kk * RB), *reinterpret_cast<const float4*>(up + kk * GC));
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
                *reinterpret_cast<float4*>(part + ((warp * 64) + a_ct * 4 + jj) * RB + a_rt * 4) = make_float4(acc[0][jj], acc[1][jj], acc[2][jj], acc[3][jj]);
        }
        __syncthreads();
        float az = 0.f, ar = 0.f;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) { az += part[(ks * 64 + eu) * RB + er]; ar += part[(ks * 64 + UPC + eu) * RB + er]; }
        const float hown = hcur[j * RB + er];
        const float z = hard_sigmoid(xz + az);
        const float r = hard_sigmoid(xr + ar);
        if (tid == 0) bar_expect(&barR[cur], XCHG_BYTES);
        push4_async(rhc, &barR[cur], j * RB + er, r * hown, er);
        bar_wait(&barR[cur], ph);
        // ---- phase B: candidate pre-activation: partB[ks][col 0..31][row]  (partB was last read before this step's first barrier)
        {
            float acc[4][4] = {};
            const float* hp = rhc + (b_ks * 16) * RB + b_rt * 4;
            const float* up = Us + (b_ks * 16) * GC + 2 * UPC + b_ct * 4;
#pragma unroll 8
            for (int kk = 0; kk < 16; ++kk)
                fma44(acc, *reinterpret_cast<const float4*>(hp + kk * RB), *reinterpret_cast<const float4*>(up + kk * GC));
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
                *reinterpret_cast<float4*>(partB + ((b_ks * 32) + b_ct * 4 + jj) * RB + b_rt * 4) = make_float4(acc[0][jj], acc[1][jj], acc[2][jj], acc[3][jj]);
        }
        __syncthreads();
        float ah = 0.f;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) ah += partB[(ks * 32 + eu) * RB + er];
        const float hh = tanhf(xh + ah);
        const float hn = z * hown + (1.f - z) * hh;
        if (tid == 0) bar_expect(&barH[cur], XCHG_BYTES);
        push4_async(hnxt, &barH[cur], j * RB + er, hn, er);
        if (valid) {
            const size_t o = ((size_t)b * T + t) * 2 + dir;
            hs[o * U + j] = hn;
            if (gates) { float* g = gates + o * (3 * U); g[j] = z; g[U + j] = r; g[2 * U + j] = hh; }
        }
        xz = nxz; xr = nxr; xh = nxh;
        // no barrier here: the next step writes `part` (last read before this step's second barrier), not partB
    }
    bar_wait(&barH[(T - 1) & 1], ((T - 1) >> 1) & 1);       // drain: all stores targeting this CTA have landed
    cluster.sync();                                         // nobody exits while a peer may still address its shared memory
}

// ------------------------------------------------------------------------------------------------- backward (BPTT)
// smem: UT[96][256] | da[3][32][8] | recvA[8][32][8] | recvB[8][32][8] | half[2048]
constexpr int BWD_SMEM_FLOATS = GC * U + 3 * UPC * RB + 2 * NCTA * UPC * RB + 2048 + 8;   // + 2 mbarriers

// partial[r][k] = sum_{u in [u0,u1)} da[gsel][u][r] * UT[col0+u][k]   for this thread's 4 rows x 4 k, accumulated into acc
__device__ __forceinline__ void bwd_dot(float (&acc)[4][4], const float* __restrict__ da_g, const float* __restrict__ UTg, int u0, int u1, int rg, int kq)
{
#pragma unroll 8
    for (int u = u0; u < u1; ++u)
        fma44(acc, *reinterpret_cast<const float4*>(da_g + u * RB + rg * 4), *reinterpret_cast<const float4*>(UTg + (size_t)u * U + kq * 4));
}

// add the other u-half's partial and push the 4(row) x 4(k) tile as four 16-byte st.async row vectors into the owner CTA's receive
// buffer (slot of this source CTA), each signalling 16 bytes on the owner's mbarrier
__device__ __forceinline__ void push_tile_async(float (&acc)[4][4], const float* __restrict__ half, float* recv_local, uint64_t* bar_local,
                                                int owner, int crank, int rg, int kq, int u0)
{
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 hv = *reinterpret_cast<const float4*>(half + (rg * 4 + i) * U + kq * 4);
        acc[i][0] += hv.x; acc[i][1] += hv.y; acc[i][2] += hv.z; acc[i][3] += hv.w;
    }
    const uint32_t rbar = map_rank(smem_addr(bar_local), owner);
#pragma unroll
    for (int q = 0; q < 4; ++q)
        st_async_v4(map_rank(smem_addr(recv_local + crank * (UPC * RB) + (u0 + q) * RB + rg * 4), owner),
                    make_float4(acc[0][q], acc[1][q], acc[2][q], acc[3][q]), rbar);
}

__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(256, 1)
gru_bwd_cluster_kernel(const float* __restrict__ dout, const float* __restrict__ hs, const float* __restrict__ gates,
                       const float* __restrict__ U0, const float* __restrict__ U1,
                       float* __restrict__ dxp, float* __restrict__ hprev_out, float* __restrict__ rh_out, int B, int T)
{
    extern __shared__ __align__(16) float sm[];
    float* UT = sm;                                 // [96][256]: UT[g*32+u][k] = U[k][g*256 + crank*32 + u]
    float* da = UT + GC * U;                        // [3][32][8]  (gate, unit, row)
    float* recvA = da + 3 * UPC * RB;               // [8 src][32][8]
    float* recvB = recvA + NCTA * UPC * RB;
    float* half = recvB + NCTA * UPC * RB;          // [8 rows][256 k] second u-half partial
    uint64_t* barA = reinterpret_cast<uint64_t*>(half + 2048);   // d(r*h) reduce-scatter
    uint64_t* barB = barA + 1;                                    // dh_{t-1} reduce-scatter
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int dir = blockIdx.y, b0 = (blockIdx.x / NCTA) * RB;
    const float* Um = dir ? U1 : U0;
    const int tid = threadIdx.x;
    // transposed U shard -> shared memory: lanes run along k (conflict-free transposed stores), 16-byte loads along the unit axis,
    // 8 independent loads in flight per thread
#pragma unroll
    for (int it = 0; it < 24; it += 8) {
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = it + q, g = r >> 3, u4 = r & 7;                  // r = (gate, unit quad); k = tid
            v[q] = __ldg(reinterpret_cast<const float4*>(Um + (size_t)tid * (3 * U) + g * U + crank * UPC + u4 * 4));
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c0 = (it + q) * 4;                                   // column = gate*32 + unit
            UT[(size_t)(c0 + 0) * U + tid] = v[q].x; UT[(size_t)(c0 + 1) * U + tid] = v[q].y;
            UT[(size_t)(c0 + 2) * U + tid] = v[q].z; UT[(size_t)(c0 + 3) * U + tid] = v[q].w;
        }
    }
    if (tid == 0) {
        bar_init(barA, 1); bar_init(barB, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();

    const int er = tid & 7, eu = tid >> 3;
    const int j = crank * UPC + eu;
    const int b = b0 + er;
    const bool valid = b < B;
    const int bb = valid ? b : B - 1;
    // compute identity: u-half (2) x row-group (2 x 4 rows) x k-quad (64 x 4 k)
    const int uh = tid >> 7, rg = (tid >> 6) & 1, kq = tid & 63;

    float dh = 0.f;
    // software pipeline: the loads of step s-1 are issued at the top of step s
    auto load_step = [&](int s, float& z, float& r, float& hh, float& hp, float& dov) {
        const int t = dir ? T - 1 - s : s;
        const int tp = dir ? t + 1 : t - 1;
        const size_t o = ((size_t)bb * T + t) * 2 + dir;
        const float* g = gates + o * (3 * U);
        z = __ldg(g + j); r = __ldg(g + U + j); hh = __ldg(g + 2 * U + j);
        hp = (s > 0) ? __ldg(hs + (((size_t)bb * T + tp) * 2 + dir) * U + j) : 0.f;
        dov = valid ? __ldg(dout + o * U + j) : 0.f;
    };
    float nz, nr, nhh, nhp, ndo;
    load_step(T - 1, nz, nr, nhh, nhp, ndo);
    for (int s = T - 1; s >= 0; --s) {
        const int t = dir ? T - 1 - s : s;
        const uint32_t xph = (uint32_t)((T - 1 - s) & 1);       // phase parity of this step's two exchanges
        const size_t o = ((size_t)bb * T + t) * 2 + dir;
        const float z = nz, r = nr, hh = nhh, hp = nhp;
        const float dht = ndo + dh;
        if (s > 0) load_step(s - 1, nz, nr, nhh, nhp, ndo);
        const float dz = dht * (hp - hh);
        float dhn = dht * z;
        const float da_h = dht * (1.f - z) * (1.f - hh * hh);
        da[(2 * UPC + eu) * RB + er] = da_h;
        __syncthreads();
        // ---- d(r*h)[r][k] partial over the own 32 units (gate h), all 256 k; reduce-scatter into recvA
        {
            float acc[4][4] = {};
            bwd_dot(acc, da + 2 * UPC * RB, UT + (size_t)2 * UPC * U, uh * 16, uh * 16 + 16, rg, kq);
            if (uh == 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(half + (rg * 4 + i) * U + kq * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            }
            __syncthreads();
            if (uh == 0) {
                const int owner = (kq * 4) / UPC, u0 = (kq * 4) % UPC;
                push_tile_async(acc, half, recvA, barA, owner, crank, rg, kq, u0);
            }
        }
        if (tid == 0) bar_expect(barA, XCHG_BYTES);
        bar_wait(barA, xph);
        float drh = 0.f;
#pragma unroll
        for (int c = 0; c < NCTA; ++c) drh += recvA[(c * UPC + eu) * RB + er];
        const float dr = drh * hp;
        dhn = fmaf(drh, r, dhn);
        const float da_z = (z > 0.f && z < 1.f) ? 0.2f * dz : 0.f;
        const float da_r = (r > 0.f && r < 1.f) ? 0.2f * dr : 0.f;
        da[(0 * UPC + eu) * RB + er] = da_z;
        da[(1 * UPC + eu) * RB + er] = da_r;
        __syncthreads();
        // ---- dh_{t-1} partial from the z and r gates (64 own columns); reduce-scatter into recvB
        {
            float acc[4][4] = {};
            bwd_dot(acc, da, UT, uh * 32, uh * 32 + 32, rg, kq);     // u in [0,64): gate z then gate r (contiguous in da and UT)
            if (uh == 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(half + (rg * 4 + i) * U + kq * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            }
            __syncthreads();
            if (uh == 0) {
                const int owner = (kq * 4) / UPC, u0 = (kq * 4) % UPC;
                push_tile_async(acc, half, recvB, barB, owner, crank, rg, kq, u0);
            }
        }
        if (tid == 0) bar_expect(barB, XCHG_BYTES);
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
        __syncthreads();          // recvA/recvB/half/da of this step fully consumed before the next step's local writes
    }
    cluster.sync();               // nobody exits while a peer may still address its shared memory
}
}  // namespace

int launch_gru_fwd_cluster(const float* xp, const float* U0, const float* U1, float* hs, float* gates, int B, int T, cudaStream_t st)
{
    if (B <= 0) return CRNN_OK;
    const size_t smem = sizeof(float) * FWD_SMEM_FLOATS;
    static bool configured = false;
    if (!configured) { CUDA_TRY(cudaFuncSetAttribute(gru_fwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = true; }
    dim3 grid(NCTA * ceil_div(B, RB), 2);
    gru_fwd_cluster_kernel<<<grid, 256, smem, st>>>(xp, U0, U1, hs, gates, B, T);
    LAUNCH_CHECK();
    return CRNN_OK;
}

int launch_gru_bwd_cluster(const float* dout, const float* hs, const float* gates, const float* U0, const float* U1,
                           float* dxp, float* hprev, float* rh, int B, int T, cudaStream_t st)
{
    if (B <= 0) return CRNN_OK;
    const size_t smem = sizeof(float) * BWD_SMEM_FLOATS;
    static bool configured = false;
    if (!configured) { CUDA_TRY(cudaFuncSetAttribute(gru_bwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = true; }
    dim3 grid(NCTA * ceil_div(B, RB), 2);
    gru_bwd_cluster_kernel<<<grid, 256, smem, st>>>(dout, hs, gates, U0, U1, dxp, hprev, rh, B, T);
    LAUNCH_CHECK();
    return CRNN_OK;
}
