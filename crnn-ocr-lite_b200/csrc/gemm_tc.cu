// gemm_tc.cu -- 5th-gen tensor-core (tcgen05 / TMEM) kernel for the pointwise 1x1 convolutions of the depthwise-separable
// stack (reference utils.py:47) and their input-gradient:   out[m][n] = sum_k f(X[m][k]) * Wop[n][k]
//
//   forward : X = raw depthwise output (M pixels x Cin), f = BatchNorm + ReLU6 fused on load, Wop = kernel^T (Cout x Cin)
//   dX      : X = dY (M x Cout), f = identity, Wop = kernel (Cin x Cout)
//
// fp32 fidelity on the tensor cores: every operand is split a = hi + lo (hi = RN-to-tf32(a), lo = a - hi) and the product is formed by
// TWO UMMAs per 8-wide k group (tf32 main term + one bf16 UMMA carrying both cross terms, see TC_IDESC_BF16); the dW kernel still uses
// the three-term 3xTF32 form.  Either keeps the stack inside the fp32 tolerance of the parity tests (tools/two_mma_error_model.py).
//
// The problem is computed TRANSPOSED on the tensor core:  D[n][m] (128 channels = TMEM lanes, 128 pixels = TMEM columns)
//   UMMA A operand (128 x K, K-major) = 16 KB pre-swizzled weight images (hi, x) written once per step by prep_weight_images_kernel,
//                                       streamed by cp.async.bulk (TMA engine) from a dedicated loader warp;
//   UMMA B operand (128 pixels x K, K-major) = the activation tile: 8 producer warps fetch it with cp.async into a raw ring, apply
//                                       BN+ReLU6, split hi/lo and store it into the canonical SWIZZLE_128B layout.
// so the epilogue is free of transposes: a warp owns 32 consecutive channels (its TMEM lane quarter), each thread reads its channel's
// pixel values with tcgen05.ld and (a) stores are 128-byte coalesced per pixel row, (b) the BatchNorm batch statistics of the output
// (sum, sum of squares per channel) -- or the reduction pass of the following BN backward -- are plain in-thread reductions.
// Persistent CTAs (one per SM) walk a static tile schedule; two TMEM accumulator buffers overlap a tile's epilogue with the next tile's
// MMAs (xw_gemm_tc_v2_kernel below).  The MMA warp issues through the "lean" path (umma_kblock): before it, the issuing thread itself
// (descriptor assembly + per-instruction waterfall loops, ~1400 cycles per 512 cycles of tensor work) paced the kernel -- ncu r1n.
#include <stdlib.h>

#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int TC_BP = 128;        // pixels per CTA  (UMMA N)
constexpr int TC_BC = 128;        // channels per CTA (UMMA M)
constexpr int TC_BK = 32;         // fp32 per k-block = one 128-byte swizzle row
constexpr int TC_TILE_FLOATS = 128 * TC_BK;            // 4096 floats = 16 KB

// ---- canonical K-major SWIZZLE_128B placement of element (row r, k) inside a 128 x 32 fp32 tile (float offset)
__host__ __device__ __forceinline__ int sw128_off(int r, int k) {
    return (r >> 3) * 256 + (r & 7) * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded spin: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 4000000000ll) __trap();   // ~2 s
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)); return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// x = hi + lo with hi = x rounded to TF32 (nearest, ties away: the value cvt.rna.tf32.f32 returns for finite x) and lo exact in fp32.
// Two integer ops instead of the four-instruction sequence ptxas emits for the cvt (which also handles Inf/NaN; activations are finite).
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = x - hi;
}


// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=64) | [46,48) version=1 | [61,64) layout=2
// (assembled as {low word, constant high word} by umma_desc_lo / TC_DESC_HI below)
// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), K-major both, N>>3 at 17, M>>4 at 24
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BP >> 3) << 17) | ((uint32_t)(TC_BC >> 4) << 24);
// fp32-faithful product in TWO UMMAs per 8-wide k group instead of the three of 3xTF32:
//   w.x = w_hi.x_hi (kind::tf32, K = 8)  +  [ w_hi'.x_lo + w_lo.x_hi' ] (ONE kind::f16 bf16 UMMA, K = 16)
// The cross terms are 2^-11 of the main term, so bf16 operands suffice for them (error <= ~2^-19 of a product).  They share one UMMA by
// interleaving along K: the "x" tiles hold, per (row, k), one 32-bit word {bf16 slot 2k, bf16 slot 2k+1} = {w_hi', w_lo} for the weights
// and {x_lo, x_hi'} for the activations -- same bytes per row, same SWIZZLE_128B geometry and the same 32-byte K advance as the tf32 tile.
constexpr uint32_t TC_IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BP >> 3) << 17) | ((uint32_t)(TC_BC >> 4) << 24);
// {slot0 -> bits 15:0, slot1 -> bits 31:16}, round to nearest
__device__ __forceinline__ float pack_bf16x2(float slot0, float slot1) {
    uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(slot1), "f"(slot0)); return __uint_as_float(r);
}

// ---- lean MMA issue path (v2 kernel).  ncu r1n showed the single MMA-issuing thread, not shared memory, pacing the kernel: ~170 SASS
// instructions per 8 UMMAs (descriptor re-assembly on the uniform datapath, one ELECT / R2UR.BROADCAST "waterfall" loop per instruction
// because the operands were not provably warp-uniform) = ~1400 cycles per (k-block, channel sub-tile) against 512 cycles of tensor work.
// Here the whole warp executes the issue code (uniform values), ONE elect.sync per k-block picks the issuing lane, the 64-bit descriptors
// are {low word, constant high word} pairs whose low words advance by plain adds (smem addresses < 256 KB: no carry out of the 14-bit field).
constexpr uint32_t TC_DESC_HI = 64u | (1u << 14) | (2u << 29);      // SBO = 1024 B >> 4, version 1, SWIZZLE_128B  (bits 32..63 of umma_desc)
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
// the four k-groups (8 fp32 each) of one 32-wide k-block: per group one bf16 cross-term UMMA + one tf32 main-term UMMA.
//   wl / xl: descriptor low words of the stage's Whi / Xhi tiles (the x / lo tiles follow 16 KB = 1024 descriptor units later)
//   first: 0 -> the very first UMMA overwrites the accumulator (start of a tile's k range)
__device__ __forceinline__ void umma_kblock(uint32_t tacc, uint32_t wl, uint32_t xl, uint32_t first_acc, uint32_t idesc_tf32, uint32_t idesc_bf16) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 a, b;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %3, 0;\n\t"
        "add.u32 a, %1, 1024;\n\tadd.u32 b, %2, 1024;\n\tmov.b64 da, {a, %4};\n\tmov.b64 db, {b, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, p;\n\t"
        "mov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, 1;\n\t"
        "add.u32 a, %1, 1026;\n\tadd.u32 b, %2, 1026;\n\tmov.b64 da, {a, %4};\n\tmov.b64 db, {b, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, 1;\n\t"
        "add.u32 a, %1, 2;\n\tadd.u32 b, %2, 2;\n\tmov.b64 da, {a, %4};\n\tmov.b64 db, {b, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, 1;\n\t"
        "add.u32 a, %1, 1028;\n\tadd.u32 b, %2, 1028;\n\tmov.b64 da, {a, %4};\n\tmov.b64 db, {b, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, 1;\n\t"
        "add.u32 a, %1, 4;\n\tadd.u32 b, %2, 4;\n\tmov.b64 da, {a, %4};\n\tmov.b64 db, {b, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, 1;\n\t"
        "add.u32 a, %1, 1030;\n\tadd.u32 b, %2, 1030;\n\tmov.b64 da, {a, %4};\n\tmov.b64 db, {b, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, 1;\n\t"
        "add.u32 a, %1, 6;\n\tadd.u32 b, %2, 6;\n\tmov.b64 da, {a, %4};\n\tmov.b64 db, {b, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, 1;\n\t}"
        ::"r"(tacc), "r"(wl), "r"(xl), "r"(first_acc), "r"(TC_DESC_HI), "r"(idesc_tf32), "r"(idesc_bf16) : "memory");
}
// The same k-block issued for a CTA PAIR (cta_group::2, UMMA M = 256 channels x N = 256 pixels): each CTA of the pair holds its 128 weight
// rows (A) and its 128 pixel rows (B) at the SAME shared-memory offsets, the leader CTA issues, both tensor cores run, each CTA's TMEM
// receives its 128 channels x all 256 pixels.  Per MAC the pair streams half the weight bytes and reads half the B bytes a lone CTA does.
__device__ __forceinline__ void umma_kblock_pair(uint32_t tacc, uint32_t wl, uint32_t xl, uint32_t first_acc, uint32_t idesc_tf32, uint32_t idesc_bf16) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 a, b;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %3, 0;\n\t"
        "add.u32 a, %1, 1024;\n\tadd.u32 b, %2, 1024;\n\tmov.b64 da, {a, %4};\n\tmov.b64 db, {b, %4};\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %6, p;\n\t"
        "mov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
        "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, 1;\n\t"
        "add.u32 a, %1, 1026;\n\tadd.u32 b, %2, 1026;\n\tmov.b64 da, {a, %4};\n\tmov.b64 db, {b, %4};\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %6, 1;\n\t"
        "add.u32 a, %1, 2;\n\tadd.u32 b, %2, 2;\n\tmov.b64 da, {a, %4};\n\tmov.b64 db, {b, %4};\n\t"
        "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, 1;\n\t"
        "add.u32 a, %1, 1028;\n\tadd.u32 b, %2, 1028;\n\tmov.b64 da, {a, %4};\n\tmov.b64 db, {b, %4};\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %6, 1;\n\t"
        "add.u32 a, %1, 4;\n\tadd.u32 b, %2, 4;\n\tmov.b64 da, {a, %4};\n\tmov.b64 db, {b, %4};\n\t"
        "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, 1;\n\t"
        "add.u32 a, %1, 1030;\n\tadd.u32 b, %2, 1030;\n\tmov.b64 da, {a, %4};\n\tmov.b64 db, {b, %4};\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %6, 1;\n\t"
        "add.u32 a, %1, 6;\n\tadd.u32 b, %2, 6;\n\tmov.b64 da, {a, %4};\n\tmov.b64 db, {b, %4};\n\t"
        "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, 1;\n\t}"
        ::"r"(tacc), "r"(wl), "r"(xl), "r"(first_acc), "r"(TC_DESC_HI), "r"(idesc_tf32), "r"(idesc_bf16) : "memory");
}
// pair commit: the arrival is multicast to the barrier at the same offset in BOTH CTAs (mask 0b11)
__device__ __forceinline__ void umma_commit_elect_pair(uint32_t bar0, uint32_t bar1, uint32_t bar2) {
    asm volatile(
        "{\n\t.reg .pred q, r1, r2;\n\t.reg .b16 m;\n\t"
        "mov.b16 m, 3;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.and.b32 r1, %1, 0, q;\n\tsetp.ne.and.b32 r2, %2, 0, q;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t"
        "@r1 tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%1], m;\n\t"
        "@r2 tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%2], m;\n\t}"
        ::"r"(bar0), "r"(bar1), "r"(bar2) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (release, cluster scope) on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
// wait with cluster-scope acquire: for barriers that the peer CTA arrives on
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 4000000000ll) __trap();
    }
}
// tcgen05.commit by one elected lane of a converged warp (up to three barriers; 0 = skip)
__device__ __forceinline__ void umma_commit_elect(uint32_t bar0, uint32_t bar1, uint32_t bar2) {
    asm volatile(
        "{\n\t.reg .pred q, r1, r2;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.and.b32 r1, %1, 0, q;\n\tsetp.ne.and.b32 r2, %2, 0, q;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "@r1 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%1];\n\t"
        "@r2 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%2];\n\t}"
        ::"r"(bar0), "r"(bar1), "r"(bar2) : "memory");
}

struct TcArgs {
    const float* X; int ldx;          // (M, K) activations, row stride ldx
    const float* Wimg;                // weight images: [(ct*KB + kb)*2 + hl][4096]
    float* out; int ldo;              // (M, N)
    int M, N, K;
    const float* x_scale; const float* x_shift;   // optional BN+ReLU6 on X (per k)
    double* stats;                    // optional [2*N] column sum / sum of squares of out
    const float* bias; int relu; int accumulate;   // epilogue: out = [out +] relu?(acc + bias[n])
    int rev;                          // v2: walk the pixel tiles from the end of the tensor (serpentine traversal)
    // v2, optional: fused reduction pass of the BatchNorm+ReLU6 backward that consumes `out` (= dL/d relu6(bn(y))): with y = red_y[m][n]
    // (same shape / row stride as out) the epilogue accumulates stats[n] += dz, stats[N + n] += dz * xhat, dz = out * 1[0 <= y*sc+sh <= 6]
    const float* red_y; const float* red_scale; const float* red_shift; const float* red_mean; const float* red_invstd;
    BnFin fin;                        // v2, forward statistics only: BatchNorm finalize done by the last CTA (common.cuh)
};

// =====================================================================================================================
// Weight gradient of the pointwise conv:  dW[ci][co] += sum_m f(X[m][ci]) * dY[m][co]     (contraction over PIXELS)
// Both operands are "MN-major" for the tensor core (channels contiguous per pixel row), so the producers keep the natural
// 128-byte channel rows.  MN-major TF32 only exists with LayoutType SWIZZLE_128B_BASE32B (cute Layout_MN_SW128_32B_Atom,
// Swizzle<2,5,2>): atom = 4 pixel rows x 128 B, the 32-byte chunk index is XORed with the row index.  Tile =
// [mn-block of 32 channels][k-group of 4 pixels][4 rows x 128 B]; LBO = mn-block stride (4096 B), SBO = k-group stride
// (512 B); one tcgen05.mma (K = 8 tf32) consumes two k-groups (start address advances by 1024 B).  D[128 ci lanes][NB co columns] accumulates in TMEM over the CTA's pixel
// range (split-K over pixels across CTAs), epilogue = fp32 atomicAdd into the pre-zeroed gradient.
// =====================================================================================================================
template <int NB> struct DwCfg {
    static constexpr int STAGES = NB == 256 ? 2 : 3;
    static constexpr int A_FLOATS = 128 * 32;            // X tile (hi or lo): 32 pixels x 128 ci
    static constexpr int B_FLOATS = NB * 32;             // dY tile (hi or lo): 32 pixels x NB co
    static constexpr int STAGE_FLOATS = 2 * A_FLOATS + 2 * B_FLOATS;
    static constexpr int SMEM_BYTES = STAGES * STAGE_FLOATS * 4 + 1024 + 256;
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);    // cross-term UMMA of the two-instruction product (TWO = true below): D = F32, A = B = BF16, both MN-major
    static constexpr uint32_t IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
};
// MN-major SWIZZLE_128B_BASE32B descriptor: LBO = 4096 B, SBO = 512 B, version 1, layout type 1 (low word / TC_DESC_MN_HI below)
// float offset of (pixel row pl in 0..31, 16-byte chunk c16 in 0..7) inside one 32-channel mn-block of an MN-major tile
__device__ __forceinline__ int mn_off(int pl, int c16) {
    const int kr = pl & 3, kg = pl >> 2;
    return kg * 128 + kr * 32 + ((((c16 >> 1) ^ kr) << 3) | ((c16 & 1) << 2));
}
// one 32-pixel k-block of the dW kernel on the lean issue path: 4 groups of 8 pixels (two 4-row atoms = 64 descriptor units apart), per group
// the three 3xTF32 terms lo*hi, hi*lo, hi*hi.  ahi = descriptor low word of the stage's A_hi tile; A_lo follows 1024 units later, B_hi 2048,
// B_lo 2048 + BLO units.  MN-major SWIZZLE_128B_BASE32B high word: SBO = 512 B >> 4, version 1, layout type 1.
constexpr uint32_t TC_DESC_MN_HI = 32u | (1u << 14) | (1u << 29);
template <uint32_t IDESC, uint32_t BLO>
__device__ __forceinline__ void umma_kblock_mn(uint32_t tacc, uint32_t ahi, uint32_t not_first) {
#define XTY_GROUP(G, P0)                                                                                                              \
        "add.u32 a, %1, " #G "*64+1024;\n\tadd.u32 b, %1, " #G "*64+2048;\n\tmov.b64 da, {a, %3};\n\tmov.b64 db, {b, %3};\n\t"                 \
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, " P0 ";\n\t"                                                        \
        "add.u32 a, %1, " #G "*64;\n\tadd.u32 b, %1, " #G "*64+2048+%5;\n\tmov.b64 da, {a, %3};\n\tmov.b64 db, {b, %3};\n\t"                    \
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, 1;\n\t"                                                            \
        "add.u32 b, %1, " #G "*64+2048;\n\tmov.b64 db, {b, %3};\n\t"                                                                  \
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, 1;\n\t"
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 a, b;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %2, 0;\n\t"
        XTY_GROUP(0, "p") XTY_GROUP(1, "1") XTY_GROUP(2, "1") XTY_GROUP(3, "1")
        "}"
        ::"r"(tacc), "r"(ahi), "r"(not_first), "r"(TC_DESC_MN_HI), "r"(IDESC), "n"(BLO) : "memory");
#undef XTY_GROUP
}

// The same k-block as TWO UMMAs per group of 8 pixels (the product scheme of the xw kernel, TC_IDESC_BF16 above): hi*hi as kind::tf32 (K = 8
// pixels) + ONE kind::f16 bf16 UMMA (K = 16) that carries both cross terms, K-interleaved per pixel p: slot 2p = {X_hi', dY_lo}, slot 2p+1 =
// {X_lo, dY_hi'}.  The cross-term tiles take the place of the lo tiles (same bytes: 64 K-rows x 2 B instead of 32 x 4 B per channel) but are
// MN-major *16-bit* tiles, whose canonical form is the plain SWIZZLE_128B one (cute Layout_MN_SW128_Atom): [mn-block of 64 channels = 128-byte
// rows][k-group of 8 K-rows][8 rows x 128 B], 16-byte chunk index XOR row; SBO = k-group stride 1024 B (high word = TC_DESC_HI), LBO = mn-block
// stride 8192 B (low word: 512 << 16 instead of the tf32 tiles' 256 << 16).  A group of 8 pixels = 16 K-rows = 2 k-groups = 128 descriptor units.
template <uint32_t IDESC, uint32_t IDESC_BF16, uint32_t BLO>
__device__ __forceinline__ void umma_kblock_mn2(uint32_t tacc, uint32_t ahi, uint32_t not_first) {
#define XTY2_GROUP(G, P0)                                                                                                             \
        "add.u32 a, %1, " #G "*64;\n\tadd.u32 b, %1, " #G "*64+2048;\n\tmov.b64 da, {a, %3};\n\tmov.b64 db, {b, %3};\n\t"                      \
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, " P0 ";\n\t"                                                        \
        "add.u32 a, %1, " #G "*128+1024+16777216;\n\tadd.u32 b, %1, " #G "*128+2048+16777216+%7;\n\tmov.b64 da, {a, %5};\n\tmov.b64 db, {b, %5};\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, 1;\n\t"
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 a, b;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %2, 0;\n\t"
        XTY2_GROUP(0, "p") XTY2_GROUP(1, "1") XTY2_GROUP(2, "1") XTY2_GROUP(3, "1")
        "}"
        ::"r"(tacc), "r"(ahi), "r"(not_first), "r"(TC_DESC_MN_HI), "r"(IDESC), "r"(TC_DESC_HI), "r"(IDESC_BF16), "n"(BLO) : "memory");
#undef XTY2_GROUP
}
__device__ __forceinline__ void sts64(uint32_t addr, float a, float b) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}

struct DwArgs {
    const float* X; int ldx; int Cin;       // (M, Cin) raw activations (+ optional BN/ReLU6 per ci)
    const float* dY; int ldy; int Cout;     // (M, Cout)
    float* dW; int ldw;                     // (Cin, Cout), pre-zeroed / accumulated with atomics
    int M; int px_per_cta;                  // pixel range per CTA (multiple of 32)
    const float* x_scale; const float* x_shift;
    int l2_pfd;                             // k-blocks of L2 prefetch distance (0 = none; CRNN_XTY_PFD)
};

constexpr int DW_THREADS = 288;     // 8 producer warps + 1 MMA warp

// Epilogue of the dW kernels (warps 0-7, 256 threads): TMEM -> smem transpose -> coalesced red.global.add.v4.f32 into dW.
// (ncu r1d: scalar REDs straight from the TMEM layout -- lane = dW row, 1 KB apart -- were 32 sectors per instruction and ~1/3
// of the kernel; the split-K head GEMMs were pure epilogue.)  The pipeline stages are idle once acc_full fires and hold the tile.
template <int NB>
__device__ __forceinline__ void xty_epilogue(const DwArgs& a, uint64_t* accb, uint32_t tmem_base, float* stage_base, int warp, int lane, int ci0, int co0) {
    constexpr int NG = NB / 128;
    mbar_wait(accb, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    constexpr int SROW = NB + 4;                      // padded row: 16-byte aligned, conflict-free STS.128 / LDS.128
    const uint32_t S = smem_u32(stage_base);
    const int lq = warp & 3;
    const int cbeg = (warp >> 2) * (NB / 2);
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + NB / 2; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const uint32_t dst = S + (uint32_t)((lq * 32 + lane) * SROW + c0) * 4u;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            sts128(dst + j * 16, __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");    // the 8 producer/epilogue warps only (the MMA warp is not part of it)
#pragma unroll 1
    for (int i = 0; i < 16; ++i) {
        const int row = warp + 8 * i, grow = ci0 + row;
        if (grow >= a.Cin) continue;
#pragma unroll
        for (int h = 0; h < NG; ++h) {
            const int col = h * 128 + lane * 4, co = co0 + col;
            if (co < a.Cout) {
                const float4 v = lds128(S + (uint32_t)(row * SROW + col) * 4u);
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.dW + (size_t)grow * a.ldw + co), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}

template <int NB, bool TWO>
__global__ void __launch_bounds__(DW_THREADS, 1) xty_gemm_tc_kernel(DwArgs a)
{
    using C = DwCfg<NB>;
    constexpr int NG = NB / 128;            // 128-channel groups of the dY tile
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* stage_base = (float*)smem;
    uint64_t* bars = (uint64_t*)(smem + C::STAGES * C::STAGE_FLOATS * 4);
    uint64_t* full = bars; uint64_t* empty = bars + C::STAGES; uint64_t* accb = bars + 2 * C::STAGES;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * C::STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ci0 = blockIdx.x * 128, co0 = blockIdx.y * NB;
    const int p_begin = blockIdx.z * a.px_per_cta;
    int p_end = p_begin + a.px_per_cta; if (p_end > a.M) p_end = a.M;
    const int KB = (p_end - p_begin + 31) / 32;
    if (KB <= 0) { pdl_enter(); return; }      // uniform for the whole CTA

    if (tid == 0) {
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 256); mbar_init(&empty[s], 1); }
        mbar_init(accb, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(NB) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_enter();            // barriers + TMEM are set up while the previous kernel of the stream drains; no global access above this line

    if (warp < 8) {
        // ------------------------------------------------ producers (next k-block's loads stay in flight in registers)
        const int q = tid & 31;                 // channel quad inside a 128-channel group
        const int pr0 = tid >> 5;               // pixel rows pr0 + 8*i
        const int ci = ci0 + q * 4;
        const bool ci_ok = ci < a.Cin;
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.x_scale && ci_ok) { sc = __ldg(reinterpret_cast<const float4*>(a.x_scale + ci)); sh = __ldg(reinterpret_cast<const float4*>(a.x_shift + ci)); }
        const int mb = q >> 3, chunk = q & 7;
        float4 xv[4], yv[NG][4], xn[4], yn[NG][4];
        auto prefetch = [&](int kb, float4 (&xd)[4], float4 (&yd)[NG][4]) {
            const int pbase = p_begin + kb * 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int p = pbase + pr0 + 8 * i;
                xd[i] = (p < p_end && ci_ok) ? __ldg(reinterpret_cast<const float4*>(a.X + (size_t)p * a.ldx + ci)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const int co = co0 + g * 128 + q * 4;
                    yd[g][i] = (p < p_end && co < a.Cout) ? __ldg(reinterpret_cast<const float4*>(a.dY + (size_t)p * a.ldy + co)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        // Optional L2 prefetch of the k-block `l2_pfd` steps ahead (CRNN_XTY_PFD, one lane per 128-byte line; OFF).  What ncu r2y shows for this
        // kernel (block 6: 103 us, tensor pipe 63 %, 53 % of the producers' stall samples are long-scoreboard waits at the first use of the
        // registers prefetched one k-block ahead; 467 MB arrive from L2 at 4.5 TB/s = 32 GB/s per SM with 48 KB of loads in flight per SM) and what
        // did NOT move it (B200, bench.py, dW kernels per step): a third less tensor work (CRNN_XTY_2MMA) 0.839 vs 0.835 ms; TMA-fed producers
        // whose proxy fence has no loads to wait for (CRNN_XTY_TMA) 0.860 ms; this L2 prefetch at distance 2 / 3 / 5 / 8: 0.864 / 0.865 / 0.886 /
        // 0.912 ms against 0.843 ms without.  So neither the tensor pipe, nor the fence, nor DRAM latency paces it; the remaining suspect is the
        // number of outstanding L1 misses an SM can hold (1536 sectors requested per k-block and SM).
        const bool pf_lane = (q & 7) == 0;
        auto l2_prefetch = [&](int kb) {
            if (!pf_lane || kb >= KB) return;
            const int pbase = p_begin + kb * 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int p = pbase + pr0 + 8 * i;
                if (p >= p_end) continue;
                if (ci_ok) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.X + (size_t)p * a.ldx + ci));
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const int co = co0 + g * 128 + q * 4;
                    if (co < a.Cout) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.dY + (size_t)p * a.ldy + co));
                }
            }
        };
        if (a.l2_pfd > 0) for (int d = 1; d < a.l2_pfd; ++d) l2_prefetch(d);
        prefetch(0, xv, yv);
        for (int kb = 0; kb < KB; ++kb) {
            if (a.l2_pfd > 0) l2_prefetch(kb + a.l2_pfd);
            if (kb + 1 < KB) prefetch(kb + 1, xn, yn);
            const int s = kb % C::STAGES;
            const uint32_t ph = (kb / C::STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            const uint32_t ahi = smem_u32(stage_base + (size_t)s * C::STAGE_FLOATS);
            const uint32_t alo = ahi + C::A_FLOATS * 4, bhi = alo + C::A_FLOATS * 4, blo = bhi + C::B_FLOATS * 4;
            const int pbase = p_begin + kb * 32;
            const bool tail = pbase + 32 > p_end;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int pl = pr0 + 8 * i;
                float x[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
                if (a.x_scale) {
                    x[0] = relu6f(fmaf(x[0], sc.x, sh.x)); x[1] = relu6f(fmaf(x[1], sc.y, sh.y));
                    x[2] = relu6f(fmaf(x[2], sc.z, sh.z)); x[3] = relu6f(fmaf(x[3], sc.w, sh.w));
                    if ((tail && pbase + pl >= p_end) || !ci_ok) { x[0] = x[1] = x[2] = x[3] = 0.f; }
                }
                float hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split_tf32(x[e], hi[e], lo[e]);
                const uint32_t off = (uint32_t)(mb * 1024 + mn_off(pl, chunk)) * 4u;
                sts128(ahi + off, hi[0], hi[1], hi[2], hi[3]);
                // cross-term tile (TWO): K-rows 2 pl (slot 0) and 2 pl + 1 (slot 1) of this thread's 4 channels = 8 bytes each
                const int r0 = 2 * pl, kr0 = r0 & 7;
                const uint32_t offc = (uint32_t)((q >> 4) * 8192 + (r0 >> 3) * 1024 + kr0 * 128 + ((q & 1) << 3));
                const uint32_t c0 = (uint32_t)((((q & 15) >> 1) ^ kr0) << 4), c1 = (uint32_t)((((q & 15) >> 1) ^ (kr0 + 1)) << 4);
                if (TWO) {
                    sts64(alo + offc + c0, pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]));               // slot 0: X_hi'
                    sts64(alo + offc + 128u + c1, pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]));        // slot 1: X_lo
                } else {
                    sts128(alo + off, lo[0], lo[1], lo[2], lo[3]);
                }
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const float y[4] = {yv[g][i].x, yv[g][i].y, yv[g][i].z, yv[g][i].w};
                    float yh[4], yl[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) split_tf32(y[e], yh[e], yl[e]);
                    const uint32_t offb = off + (uint32_t)(g * 4) * 4096u;
                    sts128(bhi + offb, yh[0], yh[1], yh[2], yh[3]);
                    if (TWO) {
                        const uint32_t offcb = offc + (uint32_t)(g * 2) * 8192u;
                        sts64(blo + offcb + c0, pack_bf16x2(yl[0], yl[1]), pack_bf16x2(yl[2], yl[3]));          // slot 0: dY_lo
                        sts64(blo + offcb + 128u + c1, pack_bf16x2(yh[0], yh[1]), pack_bf16x2(yh[2], yh[3]));   // slot 1: dY_hi'
                    } else {
                        sts128(blo + offb, yl[0], yl[1], yl[2], yl[3]);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&full[s]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                xv[i] = xn[i];
#pragma unroll
                for (int g = 0; g < NG; ++g) yv[g][i] = yn[g][i];
            }
        }
        xty_epilogue<NB>(a, accb, tmem_base, stage_base, warp, lane, ci0, co0);
    } else {
        // ------------------------------------------------ warp 8: MMA issue (lean path: whole warp, uniform operands, one elected lane --
        // see umma_kblock; the per-instruction descriptor assembly + waterfall loops cost ~110 cycles per UMMA against 64 (NB = 128))
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t adesc0 = ((smem_u32(stage_base) >> 4) & 0x3FFFu) | (256u << 16);      // MN-major: LBO = 4096 B in the low word
        const uint32_t bar_empty = smem_u32(empty), bar_acc = smem_u32(accb);
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % C::STAGES;
            const uint32_t ph = (kb / C::STAGES) & 1;
            mbar_wait(&full[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ahi = adesc0 + (uint32_t)s * (uint32_t)(C::STAGE_FLOATS * 4 / 16);
            if (TWO) umma_kblock_mn2<C::IDESC, C::IDESC_BF16, C::B_FLOATS * 4 / 16>(tb, ahi, (uint32_t)kb);
            else umma_kblock_mn<C::IDESC, C::B_FLOATS * 4 / 16>(tb, ahi, (uint32_t)kb);
            umma_commit_elect(bar_empty + 8u * s, kb == KB - 1 ? bar_acc : 0u, 0u);
        }
    }
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(NB) : "memory");
    }
}

// =====================================================================================================================
// dW kernel, TMA-fed (round 2, CRNN_XTY_TMA=1; exact, measured 3 % SLOWER than the kernel below, so off): same tiles, descriptors, MMA issue and epilogue as xty_gemm_tc_kernel below, different PRODUCERS.
// ncu r2z showed the register-prefetching producers of that kernel paced by global-load latency: their generic->async proxy fence
// (MEMBAR.ALL.CTA) waits for the NEXT k-block's loads they have in flight, so every k-block exposes a full memory latency (3300 cycles per
// k-block against 1536 of tensor work; two UMMAs instead of three changed nothing, CRNN_XTY_2MMA).  Here the loads are the TMA unit's:
//   warp 9  (one lane)  per k-block: dY box {32 channels, 32 pixels} x NB/32 with SWIZZLE_128B_ATOM_32B straight into the stage's B_hi tile
//                       (the tensor-map swizzle IS the MN-major SWIZZLE_128B_BASE32B operand layout: 32-byte chunk XOR row-in-4; the tensor
//                       core reads the top 19 bits of each fp32 word, i.e. hi = truncation to tf32), and the raw X box {128, 32} into a ring;
//   warps 0-7           X: raw ring -> BatchNorm + ReLU6 -> {hi, lo} operand tiles;  dY: read the landed B_hi tile in place and store only
//                       lo = y - trunc_tf32(y) -- no global loads in flight in these warps, so their proxy fence costs nothing;
//   warp 8              3xTF32 issue (umma_kblock_mn) as before.
// Rows >= M and channels >= Cin / Cout arrive as zeros from the TMA unit.  2 operand stages (96 KB each at NB = 256) + 2 raw X entries = 224 KB.
// =====================================================================================================================
constexpr int DW2_THREADS = 320;    // warps 0-7 transform (+ epilogue), warp 8 MMA issue, warp 9 TMA
template <int NB> struct Dw2Cfg {
    static constexpr int STAGES = 2;
    static constexpr int RAWX = NB == 256 ? 2 : 3;
    static constexpr int A_BYTES = 128 * 32 * 4, B_BYTES = NB * 32 * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + RAWX * A_BYTES + 1024 + 256;
};
__device__ __forceinline__ void tma_box_2d(uint32_t dst_smem, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
template <int NB>
__global__ void __launch_bounds__(DW2_THREADS, 1) xty_gemm_tma_kernel(DwArgs a, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY)
{
    using C = Dw2Cfg<NB>;
    using C1 = DwCfg<NB>;
    constexpr int NG = NB / 128;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* stage_base = (float*)smem;
    unsigned char* rawx = smem + C::STAGES * C::STAGE_BYTES;
    uint64_t* bars = (uint64_t*)(rawx + C::RAWX * C::A_BYTES);
    uint64_t* full = bars; uint64_t* empty = full + C::STAGES; uint64_t* bfull = empty + C::STAGES;
    uint64_t* rfull = bfull + C::STAGES; uint64_t* rempty = rfull + C::RAWX; uint64_t* accb = rempty + C::RAWX;
    uint32_t* tmem_slot = (uint32_t*)(accb + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ci0 = blockIdx.x * 128, co0 = blockIdx.y * NB;
    const int p_begin = blockIdx.z * a.px_per_cta;
    int p_end = p_begin + a.px_per_cta; if (p_end > a.M) p_end = a.M;
    const int KB = (p_end - p_begin + 31) / 32;
    if (KB <= 0) { pdl_enter(); return; }      // uniform for the whole CTA

    if (tid == 0) {
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 8); mbar_init(&empty[s], 1); mbar_init(&bfull[s], 1); }
        for (int r = 0; r < C::RAWX; ++r) { mbar_init(&rfull[r], 1); mbar_init(&rempty[r], 8); }
        mbar_init(accb, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(NB) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_enter();

    if (warp < 8) {
        // ------------------------------------------------ transform warps: a warp = one pixel row x 128 channels per step (q = channel quad)
        const int q = tid & 31, pr0 = tid >> 5;
        const int ci = ci0 + q * 4;
        const bool ci_ok = ci < a.Cin;
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.x_scale && ci_ok) { sc = __ldg(reinterpret_cast<const float4*>(a.x_scale + ci)); sh = __ldg(reinterpret_cast<const float4*>(a.x_shift + ci)); }
        const int mb = q >> 3, chunk = q & 7;
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % C::STAGES, r = kb % C::RAWX;
            const uint32_t ahi = smem_u32(smem + (size_t)s * C::STAGE_BYTES);
            const uint32_t alo = ahi + C::A_BYTES, bhi = alo + C::A_BYTES, blo = bhi + C::B_BYTES;
            const uint32_t rx = smem_u32(rawx + (size_t)r * C::A_BYTES);
            const int pbase = p_begin + kb * 32;
            mbar_wait(&rfull[r], (kb / C::RAWX) & 1);
            mbar_wait(&bfull[s], (kb / C::STAGES) & 1);          // implies empty[s]: the loader waited for it before it issued this box
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int pl = pr0 + 8 * i;
                const float4 xr = lds128(rx + (uint32_t)(pl * 128 + q * 4) * 4u);
                float x[4] = {xr.x, xr.y, xr.z, xr.w};
                if (a.x_scale) {
                    x[0] = relu6f(fmaf(x[0], sc.x, sh.x)); x[1] = relu6f(fmaf(x[1], sc.y, sh.y));
                    x[2] = relu6f(fmaf(x[2], sc.z, sh.z)); x[3] = relu6f(fmaf(x[3], sc.w, sh.w));
                    if (pbase + pl >= p_end || !ci_ok) { x[0] = x[1] = x[2] = x[3] = 0.f; }
                }
                float hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split_tf32(x[e], hi[e], lo[e]);
                const uint32_t off = (uint32_t)(mb * 1024 + mn_off(pl, chunk)) * 4u;
                sts128(ahi + off, hi[0], hi[1], hi[2], hi[3]);
                sts128(alo + off, lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const uint32_t offb = off + (uint32_t)(g * 4) * 4096u;
                    const float4 y = lds128(bhi + offb);
                    sts128(blo + offb, y.x - __uint_as_float(__float_as_uint(y.x) & 0xffffe000u), y.y - __uint_as_float(__float_as_uint(y.y) & 0xffffe000u),
                                       y.z - __uint_as_float(__float_as_uint(y.z) & 0xffffe000u), y.w - __uint_as_float(__float_as_uint(y.w) & 0xffffe000u));
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) { mbar_arrive(&full[s]); mbar_arrive(&rempty[r]); }
        }
        xty_epilogue<NB>(a, accb, tmem_base, stage_base, warp, lane, ci0, co0);
    } else if (warp == 8) {
        // ------------------------------------------------ MMA issue (lean path, see umma_kblock)
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t adesc0 = ((smem_u32(stage_base) >> 4) & 0x3FFFu) | (256u << 16);      // MN-major: LBO = 4096 B in the low word
        const uint32_t bar_empty = smem_u32(empty), bar_acc = smem_u32(accb);
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % C::STAGES;
            mbar_wait(&bfull[s], (kb / C::STAGES) & 1);      // the TMA-written B_hi tile (long complete: the transform warps waited for it)
            mbar_wait(&full[s], (kb / C::STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ahi = adesc0 + (uint32_t)s * (uint32_t)(C::STAGE_BYTES / 16);
            umma_kblock_mn<C1::IDESC, C::B_BYTES / 16>(tb, ahi, (uint32_t)kb);
            umma_commit_elect(bar_empty + 8u * s, kb == KB - 1 ? bar_acc : 0u, 0u);
        }
    } else if (lane == 0) {
        // ------------------------------------------------ TMA loader
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % C::STAGES, r = kb % C::RAWX;
            const int p0 = p_begin + kb * 32;
            const uint32_t bhi = smem_u32(smem + (size_t)s * C::STAGE_BYTES) + 2u * C::A_BYTES;
            mbar_wait(&empty[s], ((kb / C::STAGES) & 1) ^ 1);
            mbar_arrive_expect_tx(&bfull[s], (uint32_t)C::B_BYTES);
#pragma unroll
            for (int j = 0; j < NB / 32; ++j) tma_box_2d(bhi + (uint32_t)j * 4096u, &tmY, co0 + j * 32, p0, smem_u32(&bfull[s]));
            mbar_wait(&rempty[r], ((kb / C::RAWX) & 1) ^ 1);
            mbar_arrive_expect_tx(&rfull[r], (uint32_t)C::A_BYTES);
            tma_box_2d(smem_u32(rawx + (size_t)r * C::A_BYTES), &tmX, ci0, p0, smem_u32(&rfull[r]));
        }
    }
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(NB) : "memory");
    }
}

// =====================================================================================================================
// v2 of the kernel above: PERSISTENT CTAs (one per SM) walking a static tile schedule (channel tile fastest, so the CTAs
// that run concurrently share the activation tile in L2), with
//   * producers (warps 0-7) that stream the raw fp32 activations into a 3-deep shared-memory ring with cp.async (LDGSTS: 48 KB
//     in flight per SM, no registers) and transform ring entries into the hi/lo SWIZZLE_128B tiles (the ncu capture of v1 showed
//     1.7 long-scoreboard stalls per issue, 8 % resident warps and 12 % DRAM throughput: far too few bytes in flight);
//   * two TMEM accumulator buffers (2 x 128 columns) and dedicated epilogue warps (9-16), so the TMEM->HBM epilogue of tile i
//     overlaps the MMAs of tile i+1;  barriers: full/empty per smem stage, tmem_full/tmem_empty per accumulator buffer.
// =====================================================================================================================
constexpr int TC2_PROD_WARPS = 8;                // activation-transform warps 0..7
constexpr int TC2_MMA_WARP = TC2_PROD_WARPS;     // warp 8 issues the MMAs (and owns the TMEM allocation)
constexpr int TC2_EPI_WARP0 = TC2_MMA_WARP + 1;  // epilogue warps 9..16 (two per TMEM lane quarter)
constexpr int TC2_LOAD_WARP = TC2_EPI_WARP0 + 8; // warp 17 streams the weight images (1-D bulk copies)
constexpr int TC2_XLOAD_WARP = TC2_LOAD_WARP + 1; // warp 18 streams the raw activation tiles (2-D tensor-map TMA)
constexpr int TC2_THREADS = (TC2_XLOAD_WARP + 1) * 32;
constexpr int TC2_XSTAGES = 2;                  // {Xhi, Xlo} tiles consumed by the tensor core
constexpr int TC2_WRING = 3;                    // {Whi, Wlo} weight tiles, streamed by the loader warp ahead of the MMA
constexpr int TC2_RAW = 3;                      // raw fp32 activation ring filled by the tensor-map TMA loads
// Stage budget (227 KB).  Measured alternatives (gemm_bench, block 6 forward / dX, us): 2 X stages + 3 raw entries 86 / 83 (kept);
// 3 X stages + 2 raw entries 91 / 85 -- the k-block period (~2000 cycles for 1024 cycles of tensor work) is NOT a handshake latency that a
// deeper X ring would hide; it matches the shared-memory traffic model (per k-block and CTA, in 128-byte port cycles: UMMA operand reads 1024,
// weight-tile TMA writes 512, transformed-tile stores 512, raw-ring write + read 256).
constexpr int TC2_SMEM_BYTES = (TC2_WRING * 2 + TC2_XSTAGES * 2 + TC2_RAW) * TC_TILE_FLOATS * 4 + 1024 + 256;

// 2-D tiled TMA load: box {32 fp32 along k, 128 pixel rows} of X at (k0, m0) -> dense [128][32] fp32 tile; rows >= M arrive as zeros
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// store + BN-backward reduction: dz = val * 1[0 <= y*sc+sh <= 6]; s1 += dz; s2 += dz * (y*xa + xb)
template <int LDO>
__device__ __forceinline__ void epi_store_red(float* dst, const float* __restrict__ ysrc, const uint32_t (&r)[32], int ldo_rt,
                                              float sc, float sh, float xa, float xb, float& s1, float& s2) {
    float y[32];
#pragma unroll
    for (int p = 0; p < 32; ++p) y[p] = __ldg(LDO > 0 ? ysrc + p * LDO : ysrc + (size_t)p * ldo_rt);
#pragma unroll
    for (int p = 0; p < 32; ++p) {
        const float val = __uint_as_float(r[p]);
        if (LDO > 0) dst[p * LDO] = val; else dst[(size_t)p * ldo_rt] = val;
        const float z = fmaf(y[p], sc, sh);
        const float dz = (z >= 0.f && z <= 6.f) ? val : 0.f;
        s1 += dz; s2 = fmaf(dz, fmaf(y[p], xa, xb), s2);
    }
}
template <int LDO, bool STATS>
__device__ __forceinline__ void epi_store(float* dst, const uint32_t (&r)[32], int ldo_rt, float& s1, float& s2) {
#pragma unroll
    for (int p = 0; p < 32; ++p) {
        const float val = __uint_as_float(r[p]);
        if (LDO > 0) dst[p * LDO] = val; else dst[(size_t)p * ldo_rt] = val;
        if (STATS) { s1 += val; s2 = fmaf(val, val, s2); }
    }
}

// KS > 1: split-K.  Work item t = (CTA tile t / KS, k-slice t % KS); every slice contracts KBS k-blocks and stores its partial tile
// into its own copy of the output (out + slice * split_stride) -- the caller sums the copies in a fixed order (deterministic, no atomics).
// Used for the head GEMMs whose 128-pixel x 128-channel tiling yields only 33-66 tiles for 148 SMs (dense1: M=4224, N=128, K=4608).
// PAIR = true: launched in clusters of two CTAs; a CTA pair computes 256 channels x 256 pixels per tile with cta_group::2 UMMAs (see
// umma_kblock_pair).  Every role below keeps working on "its" 128 channel rows / 128 pixel rows; only the MMA warp differs: the leader
// (cluster rank 0) issues for both CTAs once BOTH have their operand stages ready, the peer's MMA warp is a relay that forwards its CTA's
// x_full / w_full / tmem_empty events to the leader with one remote arrive each.  Stage / accumulator releases come back by multicast commit.
// NTP = channel PAIR groups (N / 256), MT = pixel tiles of 256, NSUB = KS = 1, BP = 128.
template <bool PAIR>
__global__ void __launch_bounds__(TC2_THREADS, 1) xw_gemm_tc_v2_kernel(TcArgs a, const __grid_constant__ CUtensorMap tmx, int NTP, int MT, int NSUB, int KS, long long split_stride, int BP)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* w_base = (float*)smem;                                               // [WRING][Whi | Wlo]
    float* x_base = w_base + (size_t)TC2_WRING * 2 * TC_TILE_FLOATS;            // [XSTAGES][Xhi | Xlo]
    float* raw_base = x_base + (size_t)TC2_XSTAGES * 2 * TC_TILE_FLOATS;        // [RAW][128 x 32 fp32]
    uint64_t* bars = (uint64_t*)(raw_base + (size_t)TC2_RAW * TC_TILE_FLOATS);
    uint64_t* wfull = bars; uint64_t* wempty = wfull + TC2_WRING;
    uint64_t* xfull = wempty + TC2_WRING; uint64_t* xempty = xfull + TC2_XSTAGES;
    uint64_t* tfull = xempty + TC2_XSTAGES; uint64_t* tempty = tfull + 2;
    uint64_t* rawfull = tempty + 2; uint64_t* rawempty = rawfull + TC2_RAW;
    uint64_t* pxfull = rawempty + TC2_RAW; uint64_t* pwfull = pxfull + TC2_XSTAGES; uint64_t* ptempty = pwfull + TC2_WRING;   // PAIR, leader: the peer's events
    uint32_t* tmem_slot = (uint32_t*)(ptempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KB = a.K / TC_BK;
    const int KBS = (KB + KS - 1) / KS;                 // k-blocks per slice
    const int total = NTP * MT * KS;   // CTA tile = 128 pixels x (NSUB x 128) channels: the transformed X stage feeds NSUB accumulators
    auto kb_lo = [&](int t) { return (t % KS) * KBS; };
    auto kb_hi = [&](int t) { const int e = (t % KS) * KBS + KBS; return e < KB ? e : KB; };
    const int crank = PAIR ? (int)cluster_ctarank() : 0;                   // rank inside the CTA pair
    const int wid = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;       // worker = CTA, or CTA pair
    const int nwork = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    // first channel tile this CTA computes / first pixel row of the (pair) tile / first pixel row this CTA stages as the B operand
    auto tile_ct0 = [&](int t) { return PAIR ? (t % NTP) * 2 + crank : ((t / KS) % NTP) * NSUB; };
    auto tile_m0 = [&](int t) { const int mt = (t / KS) / NTP; return (a.rev ? MT - 1 - mt : mt) * (PAIR ? 2 * TC_BP : BP); };
    auto tile_m0_own = [&](int t) { return tile_m0(t) + (PAIR ? crank * TC_BP : 0); };

    if (tid == 0) {
        for (int s = 0; s < TC2_WRING; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
        for (int s = 0; s < TC2_XSTAGES; ++s) { mbar_init(&xfull[s], TC2_PROD_WARPS); mbar_init(&xempty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 256); }
        for (int s = 0; s < TC2_RAW; ++s) { mbar_init(&rawfull[s], 1); mbar_init(&rawempty[s], TC2_PROD_WARPS); }
        for (int s = 0; s < TC2_XSTAGES; ++s) mbar_init(&pxfull[s], 1);
        for (int s = 0; s < TC2_WRING; ++s) mbar_init(&pwfull[s], 1);
        for (int b = 0; b < 2; ++b) mbar_init(&ptempty[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC2_MMA_WARP) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (PAIR) cluster_sync_all();            // both CTAs' barriers are initialised before anybody arrives on a remote one
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_enter();            // barriers + TMEM are set up while the previous kernel of the stream drains; no global access above this line

    if (warp < TC2_PROD_WARPS) {
        // =========================== activation transform warps ===========================
        // The raw fp32 tile of every k-block arrives in the ring by ONE tensor-map TMA load issued by warp 18 (rows >= M zero-filled by the
        // TMA unit); these warps only read it back (each thread the same 4 (row, 16-byte chunk) slots of every entry), apply BN+ReLU6, split
        // into {hi, {lo, hi'}} and store the two SWIZZLE_128B operand tiles.  They have NO global loads of their own in flight, which is the
        // point: the generic->async proxy fence below compiles to MEMBAR.ALL.CTA, and while these warps still issued the cp.async prefetches
        // themselves (round 1) that membar waited for the prefetched k-blocks to land -- ncu r2b: long-scoreboard stalls on the fence, one
        // full memory latency per k-block, x_full the barrier the MMA warp waited on.
        const int c8 = tid & 7, r0 = tid >> 3;                                   // r0 in 0..31; rows r0 + 32 i
        const uint32_t raw_u32 = smem_u32(raw_base) + (uint32_t)(r0 * TC_BK + c8 * 4) * 4u;
        const uint32_t x_u32 = smem_u32(x_base) + (uint32_t)((r0 >> 3) * 256 + (r0 & 7) * 32 + ((c8 ^ (r0 & 7)) << 2)) * 4u;
        const bool bn = a.x_scale != nullptr;
        // BN scale / shift of this thread's 4 k-columns: fetched one k-block ahead from the (L2-resident) per-channel vectors; these small
        // loads have long landed when the proxy fence below executes
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bn) { sc = __ldg(reinterpret_cast<const float4*>(a.x_scale + c8 * 4)); sh = __ldg(reinterpret_cast<const float4*>(a.x_shift + c8 * 4)); }
        uint32_t it = 0;
        for (int t = wid; t < total; t += nwork) {
            const int m0 = tile_m0_own(t);
            const bool tail = m0 + BP > a.M;
            const int kb0 = kb_lo(t), kb1 = kb_hi(t);
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % TC2_XSTAGES;
                const uint32_t ph = (it / TC2_XSTAGES) & 1;
                const int slot = it % TC2_RAW;
                float4 scn = sc, shn = sh;                // next k-block's scale/shift (k-blocks of one tile are consecutive; a new tile restarts at kb0)
                if (bn) {
                    const int kn = (kb + 1 == kb1 ? kb0 : kb + 1) * TC_BK + c8 * 4;
                    scn = __ldg(reinterpret_cast<const float4*>(a.x_scale + kn)); shn = __ldg(reinterpret_cast<const float4*>(a.x_shift + kn));
                }
                mbar_wait(&rawfull[slot], (it / TC2_RAW) & 1);                   // the TMA load of ring entry `slot` has landed
                float4 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)      // rows >= BP (pixel tiles narrower than 128) are neither loaded by the TMA box nor read by the UMMAs
                    v[i] = (r0 + 32 * i < BP) ? lds128(raw_u32 + (uint32_t)slot * (TC_TILE_FLOATS * 4u) + (uint32_t)i * (32u * TC_BK * 4u)) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (bn) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[i].x = relu6f(fmaf(v[i].x, sc.x, sh.x)); v[i].y = relu6f(fmaf(v[i].y, sc.y, sh.y));
                        v[i].z = relu6f(fmaf(v[i].z, sc.z, sh.z)); v[i].w = relu6f(fmaf(v[i].w, sc.w, sh.w));
                    }
                    if (tail) {                           // last pixel tile only: rows >= M must stay zero after the shift
#pragma unroll
                        for (int i = 0; i < 4; ++i) if (m0 + r0 + 32 * i >= a.M) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                float4 hv[4], lv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float l0, l1, l2, l3;
                    split_tf32(v[i].x, hv[i].x, l0); split_tf32(v[i].y, hv[i].y, l1); split_tf32(v[i].z, hv[i].z, l2); split_tf32(v[i].w, hv[i].w, l3);
                    lv[i] = make_float4(pack_bf16x2(l0, hv[i].x), pack_bf16x2(l1, hv[i].y), pack_bf16x2(l2, hv[i].z), pack_bf16x2(l3, hv[i].w));   // {x_lo, x_hi'}
                }
                __syncwarp();                                                    // every lane's reads of the ring entry have been consumed above
                if (lane == 0) mbar_arrive(&rawempty[slot]);                     // -> the TMA warp may refill it
                mbar_wait(&xempty[s], ph ^ 1);
                const uint32_t xhi = x_u32 + (uint32_t)s * (2u * TC_TILE_FLOATS * 4u), xlo = xhi + TC_TILE_FLOATS * 4u;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (r0 + 32 * i >= BP) continue;
                    sts128(xhi + (uint32_t)i * 4096u, hv[i].x, hv[i].y, hv[i].z, hv[i].w);       // row r0 + 32 i: 4 eight-row groups = 4 x 1024 B further
                    sts128(xlo + (uint32_t)i * 4096u, lv[i].x, lv[i].y, lv[i].z, lv[i].w);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&xfull[s]);
                sc = scn; sh = shn;
            }
        }
    } else if (warp == TC2_XLOAD_WARP) {
        // =========================== raw activation loader: one thread, one tensor-map TMA load per k-block ===========================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmx)) : "memory");
            uint32_t it = 0;
            for (int t = wid; t < total; t += nwork) {
                const int m0 = tile_m0_own(t);
                const int kb0 = kb_lo(t), kb1 = kb_hi(t);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int slot = it % TC2_RAW;
                    mbar_wait(&rawempty[slot], ((it / TC2_RAW) & 1) ^ 1);
                    mbar_arrive_expect_tx(&rawfull[slot], (uint32_t)BP * TC_BK * 4);        // the box is BP rows x 32 fp32
                    tma_load_2d(raw_base + (size_t)slot * TC_TILE_FLOATS, &tmx, kb * TC_BK, m0, &rawfull[slot]);
                }
            }
        }
    } else if (warp == TC2_LOAD_WARP) {
        // =========================== weight loader: one thread streams the pre-swizzled hi/lo images (TMA bulk copies) ===========================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = wid; t < total; t += nwork) {
                const int ct0 = tile_ct0(t);
                const int kb0 = kb_lo(t), kb1 = kb_hi(t);
                for (int kb = kb0; kb < kb1; ++kb) {
                    for (int h = 0; h < NSUB; ++h, ++it) {
                        const int ws = it % TC2_WRING;
                        mbar_wait(&wempty[ws], ((it / TC2_WRING) & 1) ^ 1);
                        float* Whi = w_base + (size_t)ws * (2 * TC_TILE_FLOATS);
                        const float* src = a.Wimg + ((size_t)((ct0 + h) * KB + kb) * 2) * TC_TILE_FLOATS;
                        mbar_arrive_expect_tx(&wfull[ws], 2 * TC_TILE_FLOATS * 4);
                        bulk_g2s(Whi, src, 2 * TC_TILE_FLOATS * 4, &wfull[ws]);      // hi and lo images are contiguous: one 32 KB copy
                    }
                }
            }
        }
    } else if (warp == TC2_MMA_WARP) {
        // =========================== MMA issuer ===========================
        // the whole warp runs this code with warp-uniform values; one elected lane issues (umma_kblock / umma_commit_elect)
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t wdesc0 = umma_desc_lo(smem_u32(w_base)), xdesc0 = umma_desc_lo(smem_u32(x_base));
        const uint32_t bar_wempty = smem_u32(wempty), bar_xempty = smem_u32(xempty), bar_tfull = smem_u32(tfull);
        // UMMA N = pixels per tile (a multiple of 16 chosen by the launcher, see pick_pixel_tile): idesc bits [17,23) = N >> 3
        const uint32_t idesc_tf32 = (TC_IDESC & ~(0x3Fu << 17)) | ((uint32_t)(BP >> 3) << 17);
        const uint32_t idesc_bf16 = (TC_IDESC_BF16 & ~(0x3Fu << 17)) | ((uint32_t)(BP >> 3) << 17);
        uint32_t it = 0, wit = 0;
        int j = 0;
        if (PAIR && crank != 0) {
            // peer CTA of a pair: relay.  Forwards "my accumulator buffer is drained / my X stage is written / my W tile has landed" to
            // the leader, which issues the UMMAs for both CTAs.
            for (int t = wid; t < total; t += nwork, ++j) {
                const int buf = j & 1;
                mbar_wait(&tempty[buf], ((j >> 1) & 1) ^ 1);
                if (lane == 0) mbar_arrive_remote(&ptempty[buf], 0);
                for (int kb = 0; kb < KB; ++kb, ++it, ++wit) {
                    const int xs = it % TC2_XSTAGES, ws = wit % TC2_WRING;
                    mbar_wait(&xfull[xs], (it / TC2_XSTAGES) & 1);
                    if (lane == 0) mbar_arrive_remote(&pxfull[xs], 0);
                    mbar_wait(&wfull[ws], (wit / TC2_WRING) & 1);
                    if (lane == 0) mbar_arrive_remote(&pwfull[ws], 0);
                }
            }
        } else {
        const uint32_t pidesc_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);      // cta_group::2: M = N = 256
        const uint32_t pidesc_bf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
        for (int t = wid; t < total; t += nwork, ++j) {
            const int buf = j & 1;
            mbar_wait(&tempty[buf], ((j >> 1) & 1) ^ 1);          // epilogue has drained this accumulator buffer
            if (PAIR) mbar_wait_cluster(&ptempty[buf], (j >> 1) & 1);   // ... and so has the peer's (one relay arrival per use of the buffer)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int kb0 = kb_lo(t), kb1 = kb_hi(t);
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int xs = it % TC2_XSTAGES;
                mbar_wait(&xfull[xs], (it / TC2_XSTAGES) & 1);
                if (PAIR) mbar_wait_cluster(&pxfull[xs], (it / TC2_XSTAGES) & 1);
                const uint32_t xl = xdesc0 + (uint32_t)xs * (2u * TC_TILE_FLOATS * 4u / 16u);
                for (int h = 0; h < NSUB; ++h, ++wit) {
                    const int ws = wit % TC2_WRING;
                    const uint32_t tacc = tb + (uint32_t)(buf * 256 + h * 128);
                    mbar_wait(&wfull[ws], (wit / TC2_WRING) & 1);
                    if (PAIR) mbar_wait_cluster(&pwfull[ws], (wit / TC2_WRING) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t wl = wdesc0 + (uint32_t)ws * (2u * TC_TILE_FLOATS * 4u / 16u);
                    const bool last_h = h == NSUB - 1;
                    const uint32_t b1 = last_h ? bar_xempty + 8u * xs : 0u, b2 = (last_h && kb == kb1 - 1) ? bar_tfull + 8u * buf : 0u;
                    if (PAIR) {
                        umma_kblock_pair(tacc, wl, xl, (uint32_t)(kb - kb0), pidesc_tf32, pidesc_bf16);
                        umma_commit_elect_pair(bar_wempty + 8u * ws, b1, b2);
                    } else {
                        umma_kblock(tacc, wl, xl, (uint32_t)(kb - kb0), idesc_tf32, idesc_bf16);
                        umma_commit_elect(bar_wempty + 8u * ws, b1, b2);
                    }
                }
            }
        }
        }
    } else {
        // =========================== epilogue warps 9..16 ===========================
        // Two warps per TMEM lane quarter, each draining half of the 128 pixel columns.  (ncu on the 4-warp version: ~47 % of all
        // stall samples in this loop and the MMA warp waiting on tmem_empty -> the epilogue, not the tensor core, paced the tile.)
        const int q = warp & 3;                                     // TMEM lane quarter this warp may access
        const int half = (warp - TC2_EPI_WARP0) >> 2;                           // 0: columns 0..63, 1: columns 64..127
        // BN statistics: per-thread fp64 running sums over all tiles of the same channel group, flushed with two atomics per channel
        // when the group changes / at the end (per-tile atomics = ~5 k same-address fp64 atomics per channel in block 2/3: ~50 us of
        // pure L2 atomic serialisation in the DIAG=15 skeleton run)
        double S1[2] = {0.0, 0.0}, S2[2] = {0.0, 0.0};
        int cur_ct0 = -1;
        auto flush_stats = [&]() {
            if (!a.stats || cur_ct0 < 0) return;
            for (int h = 0; h < NSUB; ++h) {
                const int n = (cur_ct0 + h) * TC_BC + q * 32 + lane;
                if (n < a.N) { atomicAdd(a.stats + n, S1[h]); atomicAdd(a.stats + a.N + n, S2[h]); }
                S1[h] = S2[h] = 0.0;
            }
        };
        int j = 0;
        const int NCOL = PAIR ? 2 * TC_BP : BP;                                       // pixel columns of one accumulator
        const int HC = PAIR ? TC_BP : 64;                                             // columns per epilogue warp (two warps per lane quarter)
        for (int t = wid; t < total; t += nwork, ++j) {
            const int buf = j & 1;
            const int ct0 = tile_ct0(t), m0 = tile_m0(t);
            const int m_end = m0 + NCOL < a.M ? m0 + NCOL : a.M;                      // first pixel row NOT in this tile
            float* const outp = a.out + (size_t)(t % KS) * (size_t)split_stride;      // this k-slice's copy of the output
            if (ct0 != cur_ct0) { flush_stats(); cur_ct0 = ct0; }
            if (a.red_y) {
                // the fused reduction reads y = red_y at the tile's positions: pull this warp's 64 pixel rows x 128 B (per channel tile) into
                // L2 while the tensor core is still working on the tile (the loads sit on the epilogue's critical path otherwise)
                for (int h = 0; h < NSUB; ++h) {
                    const int nb = (ct0 + h) * TC_BC + q * 32;
                    for (int pp = 0; pp < HC; pp += 32) {
                        const int m = m0 + half * HC + pp + lane;
                        if (m < m_end && nb < a.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.red_y + (size_t)m * a.ldo + nb));
                    }
                }
            }
            mbar_wait(&tfull[buf], (j >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
          for (int h = 0; h < NSUB; ++h) {
            const int n = (ct0 + h) * TC_BC + q * 32 + lane;
            const bool n_ok = n < a.N;
            const float bias = (a.bias && n_ok) ? __ldg(a.bias + n) : 0.f;
            float rsc = 0.f, rsh = 0.f, rxa = 0.f, rxb = 0.f;                    // fused BN-backward reduction: per-channel constants
            if (a.red_y && n_ok) { rsc = __ldg(a.red_scale + n); rsh = __ldg(a.red_shift + n); rxa = __ldg(a.red_invstd + n); rxb = -__ldg(a.red_mean + n) * rxa; }
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
            for (int c0 = half * HC; c0 < half * HC + HC && c0 < NCOL; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + h * 128 + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float* dst = outp + (size_t)(m0 + c0) * a.ldo + n;      // 32 lanes = 32 consecutive channels: 128-byte coalesced rows
                if (n_ok && m0 + c0 + 32 <= m_end && !a.accumulate) {   // fast paths: full 32-pixel chunk, no per-element predicate
                    // row stride known at compile time for the conv-stack widths -> the 32 stores use immediate offsets (the generic
                    // loop costs a 64-bit add per store; the epilogue warps share the issue slots with the activation producers)
                    if (a.red_y) {                                              // dX + reduction pass of the following BN/ReLU6 backward
                        const float* ysrc = a.red_y + (size_t)(m0 + c0) * a.ldo + n;
                        switch (a.ldo) {
                            case 64: epi_store_red<64>(dst, ysrc, r, 0, rsc, rsh, rxa, rxb, s1, s2); break;
                            case 128: epi_store_red<128>(dst, ysrc, r, 0, rsc, rsh, rxa, rxb, s1, s2); break;
                            case 256: epi_store_red<256>(dst, ysrc, r, 0, rsc, rsh, rxa, rxb, s1, s2); break;
                            case 512: epi_store_red<512>(dst, ysrc, r, 0, rsc, rsh, rxa, rxb, s1, s2); break;
                            default: epi_store_red<0>(dst, ysrc, r, a.ldo, rsc, rsh, rxa, rxb, s1, s2);
                        }
                    } else if (a.stats) {                                       // conv forward: store + BN statistics
                        switch (a.ldo) {
                            case 128: epi_store<128, true>(dst, r, 0, s1, s2); break;
                            case 256: epi_store<256, true>(dst, r, 0, s1, s2); break;
                            case 512: epi_store<512, true>(dst, r, 0, s1, s2); break;
                            default: epi_store<0, true>(dst, r, a.ldo, s1, s2);
                        }
                    } else if (a.bias || a.relu) {                              // head: bias (+ ReLU)
#pragma unroll
                        for (int p = 0; p < 32; ++p) {
                            float val = __uint_as_float(r[p]) + bias;
                            if (a.relu) val = fmaxf(val, 0.f);
                            dst[(size_t)p * a.ldo] = val;
                        }
                    } else {                                                    // dX: plain store
                        switch (a.ldo) {
                            case 64: epi_store<64, false>(dst, r, 0, s1, s2); break;
                            case 128: epi_store<128, false>(dst, r, 0, s1, s2); break;
                            case 256: epi_store<256, false>(dst, r, 0, s1, s2); break;
                            case 512: epi_store<512, false>(dst, r, 0, s1, s2); break;
                            default: epi_store<0, false>(dst, r, a.ldo, s1, s2);
                        }
                    }
                } else if (n_ok) {
#pragma unroll
                    for (int p = 0; p < 32; ++p) {
                        if (m0 + c0 + p < m_end) {
                            float val = __uint_as_float(r[p]) + bias;
                            if (a.relu) val = fmaxf(val, 0.f);
                            if (a.accumulate) val += *dst;
                            *dst = val;
                            if (a.red_y) {
                                const float y = __ldg(a.red_y + (size_t)(m0 + c0 + p) * a.ldo + n);
                                const float z = fmaf(y, rsc, rsh);
                                const float dz = (z >= 0.f && z <= 6.f) ? val : 0.f;
                                s1 += dz; s2 = fmaf(dz, fmaf(y, rxa, rxb), s2);
                            } else { s1 += val; s2 = fmaf(val, val, s2); }
                        }
                        dst += a.ldo;
                    }
                }
            }
            if (h == 0) { S1[0] += (double)s1; S2[0] += (double)s2; } else { S1[1] += (double)s1; S2[1] += (double)s2; }
          }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&tempty[buf]);                               // accumulator buffer may be overwritten
        }
        flush_stats();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    bn_finalize_tail(a.fin);                 // every CTA of the grid gets here, its statistics flushed
    if (PAIR) cluster_sync_all();            // neither CTA leaves (or frees TMEM) while the other may still signal its barriers / read its tiles
    if (warp == TC2_MMA_WARP) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// Wimg[((ct*KB + kb)*2 + hl)*4096 + sw128_off(r, kk)] for Wop[n = ct*128 + r][k = kb*32 + kk]; rows n >= N are zero.
//   transposed=1: Wop[n][k] = W[k*ldw + n]  (forward: Keras kernel is (Cin, Cout));  0: Wop[n][k] = W[n*ldw + k]  (dX)
__global__ void prep_weight_images_kernel(const float* __restrict__ W, int ldw, int N, int K, int transposed, float* __restrict__ img)
{ pdl_enter();
    const int KB = K / TC_BK;
    const int NT = (N + TC_BC - 1) / TC_BC;
    const long long total = (long long)NT * TC_BC * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int n, k;
        if (transposed) { n = (int)(i % (NT * TC_BC)); k = (int)(i / (NT * TC_BC)); }    // coalesced reads along n
        else            { k = (int)(i % K); n = (int)(i / K); }                           // coalesced reads along k
        float w = 0.f;
        if (n < N) w = transposed ? W[(size_t)k * ldw + n] : W[(size_t)n * ldw + k];
        const uint32_t hi = to_tf32(w);
        const float lo = w - __uint_as_float(hi);
        const int ct = n / TC_BC, r = n % TC_BC, kb = k / TC_BK, kk = k % TC_BK;
        float* t = img + ((size_t)(ct * KB + kb) * 2) * TC_TILE_FLOATS;
        const int off = sw128_off(r, kk);
        t[off] = __uint_as_float(hi);
        t[TC_TILE_FLOATS + off] = pack_bf16x2(__uint_as_float(hi), lo);          // {w_hi', w_lo}
    }
}
// all weight images of a step in ONE launch (22 tiny kernels on the side branch before): job j covers the flat index range
// [start[j], start[j+1]) of its (channel-tile-padded N) x K element space
__global__ void prep_weight_images_batch_kernel(TcPrepBatch t)
{ pdl_enter();
    const long long total = t.start[t.n];
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        int j = 0;
        while (g >= t.start[j + 1]) ++j;
        const long long i = g - t.start[j];
        const int N = t.N[j], K = t.K[j], ldw = t.ldw[j];
        const int KB = K / TC_BK, NT = (N + TC_BC - 1) / TC_BC;
        int n, k;
        if (t.transposed[j]) { n = (int)(i % (NT * TC_BC)); k = (int)(i / (NT * TC_BC)); }
        else                 { k = (int)(i % K); n = (int)(i / K); }
        float w = 0.f;
        if (n < N) w = t.transposed[j] ? t.W[j][(size_t)k * ldw + n] : t.W[j][(size_t)n * ldw + k];
        const uint32_t hi = to_tf32(w);
        const float lo = w - __uint_as_float(hi);
        const int ct = n / TC_BC, r = n % TC_BC, kb = k / TC_BK, kk = k % TC_BK;
        float* dst = t.img[j] + ((size_t)(ct * KB + kb) * 2) * TC_TILE_FLOATS;
        const int off = sw128_off(r, kk);
        dst[off] = __uint_as_float(hi);
        dst[TC_TILE_FLOATS + off] = pack_bf16x2(__uint_as_float(hi), lo);
    }
}
}  // namespace

int tc_prep_batch_add(TcPrepBatch& t, const float* W, int ldw, int N, int K, int transposed, float* img)
{
    if (K % TC_BK) { crnn_set_error("gemm_tc: K=%d must be a multiple of %d", K, TC_BK); return CRNN_ERR_INVALID; }
    if (t.n >= TcPrepBatch::MAXJ) { crnn_set_error("gemm_tc: more than %d weight images in one batch", TcPrepBatch::MAXJ); return CRNN_ERR_INVALID; }
    const int j = t.n++;
    if (j == 0) t.start[0] = 0;
    t.W[j] = W; t.img[j] = img; t.ldw[j] = ldw; t.N[j] = N; t.K[j] = K; t.transposed[j] = transposed;
    t.start[j + 1] = t.start[j] + (long long)((N + TC_BC - 1) / TC_BC) * TC_BC * K;
    return CRNN_OK;
}
int launch_prep_weight_images_batch(const TcPrepBatch& t, cudaStream_t st)
{
    if (t.n <= 0) return CRNN_OK;
    const long long total = t.start[t.n];
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
    (void)crnn_launch(prep_weight_images_batch_kernel, blocks, 256, 0, st, t);
    LAUNCH_CHECK();
    return CRNN_OK;
}

size_t tc_weight_image_floats(int N, int K) { return (size_t)((N + TC_BC - 1) / TC_BC) * (K / TC_BK) * 2 * TC_TILE_FLOATS; }

int launch_prep_weight_images(const float* W, int ldw, int N, int K, int transposed, float* img, cudaStream_t st)
{
    if (K % TC_BK) { crnn_set_error("gemm_tc: K=%d must be a multiple of %d", K, TC_BK); return CRNN_ERR_INVALID; }
    long long total = (long long)((N + TC_BC - 1) / TC_BC) * TC_BC * K;
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
    (void)crnn_launch(prep_weight_images_kernel, blocks, 256, 0, st, W, ldw, N, K, transposed, img);
    LAUNCH_CHECK();
    return CRNN_OK;
}

int xw_gemm_tc_pick_ksplit(int M, int N, int K)
{
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const int NT = (N + TC_BC - 1) / TC_BC, MT = (M + TC_BP - 1) / TC_BP, KB = K / TC_BK;
    const int NSUB = (NT % 2 == 0 && (long long)(NT / 2) * MT >= sms) ? 2 : 1;
    const long long tiles = (long long)(NT / NSUB) * MT;
    if (tiles * 2 > sms || KB < 8) return 1;
    int ks = (int)(sms / tiles);
    if (ks > KB / 4) ks = KB / 4;
    if (ks < 1) ks = 1;
    const int kbs = (KB + ks - 1) / ks;
    return (KB + kbs - 1) / kbs;                      // no empty slices
}

int launch_xw_gemm_tc(const float* X, int ldx, const float* Wimg, float* out, int ldo, int M, int N, int K,
                      const float* x_scale, const float* x_shift, double* stats, cudaStream_t st,
                      const float* bias, int relu, int accumulate, int rev, int ksplit, long long split_stride, const TcBnRed* red)
{
    g_crnn_family = CRNN_FAM_XW_TC;
    if (M <= 0 || N <= 0) return CRNN_OK;
    if (K % TC_BK || K <= 0) { crnn_set_error("gemm_tc: K=%d must be a positive multiple of %d", K, TC_BK); return CRNN_ERR_INVALID; }
    if ((ldx % 4) || (reinterpret_cast<uintptr_t>(X) & 15) || (reinterpret_cast<uintptr_t>(Wimg) & 15)) { crnn_set_error("gemm_tc: X/Wimg must be 16-byte aligned, ldx %% 4 == 0"); return CRNN_ERR_INVALID; }
    TcArgs a; a.X = X; a.ldx = ldx; a.Wimg = Wimg; a.out = out; a.ldo = ldo; a.M = M; a.N = N; a.K = K;
    a.x_scale = x_scale; a.x_shift = x_shift; a.stats = stats; a.bias = bias; a.relu = relu; a.accumulate = accumulate; a.rev = rev;
    a.red_y = nullptr; a.red_scale = a.red_shift = a.red_mean = a.red_invstd = nullptr;
    a.fin = BnFin{}; if (stats && !red) a.fin = crnn_take_bn_fin();
    if (red) {
        if (!stats || bias || relu || accumulate || ksplit > 1) { crnn_set_error("gemm_tc: the fused BN-backward reduction needs stats and the plain-store epilogue"); return CRNN_ERR_INVALID; }
        a.red_y = red->y; a.red_scale = red->scale; a.red_shift = red->shift; a.red_mean = red->mean; a.red_invstd = red->invstd;
    }
    static int nsub_env = -1, bp_env = -1, pair_env = 0;     // A/B switches: CRNN_GEMM_NSUB=1 (one channel tile per CTA tile), CRNN_GEMM_BP=<pixels per tile, multiple of 16>, CRNN_GEMM_PAIR=1
    if (nsub_env < 0) {
        const char* f = getenv("CRNN_GEMM_NSUB"); nsub_env = f ? atoi(f) : 0;
        const char* g = getenv("CRNN_GEMM_BP"); bp_env = g ? atoi(g) : 0; if (bp_env % 16 || bp_env > 128 || bp_env < 16) bp_env = 0;
        const char* q = getenv("CRNN_GEMM_PAIR"); pair_env = (q && q[0] == '1') ? 1 : 0;
    }
    const int NT = (N + TC_BC - 1) / TC_BC;
    if (ksplit > 1 && (x_scale || stats || bias || relu || accumulate || split_stride < (long long)M * ldo - (ldo - N) || ksplit > K / TC_BK)) {
        crnn_set_error("gemm_tc: split-K needs the plain-store epilogue (no BN prologue / statistics / bias / relu / accumulate) and disjoint output copies");
        return CRNN_ERR_INVALID;
    }
    if (ksplit < 1) ksplit = 1;
    static bool configured2 = false;
    static int num_sms = 148;
    if (!configured2) {
        CUDA_TRY(cudaFuncSetAttribute(xw_gemm_tc_v2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(xw_gemm_tc_v2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES));
        int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        configured2 = true;
    }
    // CTA pairs (cta_group::2, 256 channels x 256 pixels per pair tile) whenever the output is at least 256 channels wide and there are
    // enough pair tiles to keep the pairs busy.  OFF by default (CRNN_GEMM_PAIR=1 enables it): results are exact (test_gemm_tc runs both
    // schedules), and per MAC the pair moves ~45 % less through shared memory, but it MEASURED 1.6x slower (block 6 forward 140 us against
    // 86 us; ncu r2k: tensor pipe 37 %, producers waiting on x_empty): every stage hand-over crosses the CTA boundary twice (multicast
    // commit to the peer, relay arrive back to the leader) and two X stages cannot cover that round trip.
    const bool pair = pair_env == 1 && ksplit == 1 && N % 256 == 0 && (long long)(N / 256) * ((M + 255) / 256) >= num_sms / 4;
    // two channel tiles per CTA tile when there are enough tiles to fill the SMs anyway: the BN/ReLU6 + hi/lo transform of the
    // activation tile (the instruction-issue bottleneck, ncu r1d) is then shared by 2 x 128 output channels
    const int NSUB = pair ? 1 : ((NT % 2 == 0 && (long long)(NT / 2) * ((M + TC_BP - 1) / TC_BP) >= num_sms && nsub_env != 1) ? 2 : 1);
    const int NTP = pair ? NT / 2 : NT / NSUB;
    // pixels per tile (UMMA N, any multiple of 16; CRNN_GEMM_BP for A/B runs).  The persistent CTAs walk a static schedule, so 594 tiles of
    // 128 pixels on 148 SMs (blocks 4, 6, 7 at batch 64) last 5 rounds for 4.01 rounds of work; narrower tiles pack the rounds better (112:
    // 4.6 of 5) but MEASURED slower (block 6 forward 86 -> 91 us, block 5 with 96-pixel tiles 85 -> 98 us, r2h): the weight tiles are
    // re-streamed through shared memory once per pixel tile, and that traffic -- not the tensor pipe -- bounds the kernel.  128 stays.
    const int BP = (bp_env > 0 && !pair) ? bp_env : TC_BP;
    const int MT = pair ? (M + 2 * TC_BP - 1) / (2 * TC_BP) : (M + BP - 1) / BP;
    // tensor map of X for the raw-tile TMA loads: dims {K, M} fp32, row stride ldx, box {32, BP}, no swizzle, out-of-range rows read as 0
    CUtensorMap tmx;
    {
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static EncodeFn encode = nullptr;
        if (!encode) {
            void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
                cudaGetLastError(); crnn_set_error("gemm_tc: cuTensorMapEncodeTiled is not available from this driver"); return CRNN_ERR_CUDA;
            }
            encode = reinterpret_cast<EncodeFn>(fn);
        }
        const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
        const cuuint64_t gstr[1] = {(cuuint64_t)ldx * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)BP};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { crnn_set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) for X %p, M %d, K %d, ldx %d", (int)r, (const void*)X, M, K, ldx); return CRNN_ERR_CUDA; }
    }
    if (pair) {
        const long long total = (long long)NTP * MT;
        const int pairs = (int)(total < num_sms / 2 ? total : num_sms / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(TC2_THREADS); cfg.dynamicSmemBytes = TC2_SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, xw_gemm_tc_v2_kernel<true>, a, tmx, NTP, MT, 1, 1, (long long)0, TC_BP));
    } else {
        const long long total = (long long)NTP * MT * ksplit;
        const int grid = (int)(total < num_sms ? total : num_sms);
        (void)crnn_launch(xw_gemm_tc_v2_kernel<false>, grid, TC2_THREADS, TC2_SMEM_BYTES, st, a, tmx, NTP, MT, NSUB, ksplit, split_stride, BP);
    }
    LAUNCH_CHECK();
    return CRNN_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda): 2-D fp32 tensor {dim0 contiguous, dim1 rows
// of stride ld floats}, box {box0, box1}, out-of-range elements read as 0
static int tc_encode_2d(CUtensorMap* tm, const float* base, long long dim0, long long dim1, long long ld, int box0, int box1, CUtensorMapSwizzle swz) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            cudaGetLastError(); crnn_set_error("gemm_tc: cuTensorMapEncodeTiled is not available from this driver"); return CRNN_ERR_CUDA;
        }
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { crnn_set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) for %p, dims %lld x %lld, ld %lld, box %d x %d", (int)r, (const void*)base, dim0, dim1, ld, box0, box1); return CRNN_ERR_CUDA; }
    return CRNN_OK;
}

int launch_xty_gemm_tc(const float* X, int ldx, int Cin, const float* dY, int ldy, int Cout, float* dW, int ldw, int M,
                       const float* x_scale, const float* x_shift, cudaStream_t st)
{
    g_crnn_family = CRNN_FAM_XTY_TC;
    if (M <= 0 || Cin <= 0 || Cout <= 0) return CRNN_OK;
    if ((Cin % 4) || (Cout % 4) || (ldx % 4) || (ldy % 4) || (reinterpret_cast<uintptr_t>(X) & 15) || (reinterpret_cast<uintptr_t>(dY) & 15)) {
        crnn_set_error("gemm_tc dW: channels / leading dims must be multiples of 4 and pointers 16-byte aligned"); return CRNN_ERR_INVALID;
    }
    if ((ldw % 4) || (reinterpret_cast<uintptr_t>(dW) & 15)) {
        crnn_set_error("gemm_tc dW: dW must be 16-byte aligned with ldw %% 4 == 0 (vector reductions)"); return CRNN_ERR_INVALID;
    }
    const int NB = Cout > 128 ? 256 : 128;
    const int tiles = ((Cin + 127) / 128) * ((Cout + NB - 1) / NB);
    // split-K over pixels so that tiles x splits fills the SMs in ONE wave (the kernel runs one CTA per SM: 152 CTAs on 148 SMs
    // doubled blocks 6/7's time)
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    int splits = sms / tiles; if (splits < 1) splits = 1;
    int per = ((M + splits - 1) / splits + 31) / 32 * 32;
    if (per < 32) per = 32;
    splits = (M + per - 1) / per;
    DwArgs a; a.X = X; a.ldx = ldx; a.Cin = Cin; a.dY = dY; a.ldy = ldy; a.Cout = Cout; a.dW = dW; a.ldw = ldw; a.M = M; a.px_per_cta = per;
    a.x_scale = x_scale; a.x_shift = x_shift;
    static int pfd = -1;
    if (pfd < 0) { const char* e = getenv("CRNN_XTY_PFD"); pfd = e ? atoi(e) : 0; if (pfd < 0 || pfd > 16) pfd = 0; }
    a.l2_pfd = pfd;
    dim3 grid((Cin + 127) / 128, (Cout + NB - 1) / NB, splits);
    // CRNN_XTY_TMA=1: the TMA-fed kernel (xty_gemm_tma_kernel) instead of the register-prefetching producers
    static int v1 = -1;
    if (v1 < 0) { const char* e = getenv("CRNN_XTY_TMA"); v1 = (e && e[0] == '1') ? 0 : 1; }
    if (!v1) {
        CUtensorMap tmX, tmY;
        { const int rc = tc_encode_2d(&tmX, X, Cin, M, ldx, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE); if (rc != CRNN_OK) return rc; }
        { const int rc = tc_encode_2d(&tmY, dY, Cout, M, ldy, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B); if (rc != CRNN_OK) return rc; }
        static bool conf2 = false;
        if (!conf2) {
            CUDA_TRY(cudaFuncSetAttribute(xty_gemm_tma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Dw2Cfg<256>::SMEM_BYTES));
            CUDA_TRY(cudaFuncSetAttribute(xty_gemm_tma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Dw2Cfg<128>::SMEM_BYTES));
            conf2 = true;
        }
        if (NB == 256) (void)crnn_launch(xty_gemm_tma_kernel<256>, grid, DW2_THREADS, Dw2Cfg<256>::SMEM_BYTES, st, a, tmX, tmY);
        else (void)crnn_launch(xty_gemm_tma_kernel<128>, grid, DW2_THREADS, Dw2Cfg<128>::SMEM_BYTES, st, a, tmX, tmY);
        LAUNCH_CHECK();
        return CRNN_OK;
    }
    // CRNN_XTY_2MMA=1: the two-UMMA product (tf32 main term + one bf16 UMMA with both cross terms, umma_kblock_mn2) instead of 3xTF32.  Exact to the
    // same tolerance (test_gemm_tc_dw passes either way) and a third less tensor work, but NOT faster -- measured in the step (B200, bench.py,
    // 40 steps, twice each): dW kernels 0.835 / 0.837 ms per step with 3xTF32, 0.839 / 0.839 ms with two UMMAs; step 4.986 / 4.988 vs 4.990 /
    // 4.989 ms.  The kernel is paced by its producers' operand traffic (block 6: 456 MB from L2 in 112 us = 4.1 TB/s, every X tile is read by
    // the 2 CTAs of the other output-channel tiles and every dY tile by 4), not by the tensor pipe -- so the more accurate 3xTF32 form stays.
    static int three = -1;
    if (three < 0) { const char* e = getenv("CRNN_XTY_2MMA"); three = (e && e[0] == '1') ? 0 : 1; }
    static bool configured = false;
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(xty_gemm_tc_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwCfg<256>::SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(xty_gemm_tc_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwCfg<256>::SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(xty_gemm_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwCfg<128>::SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(xty_gemm_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwCfg<128>::SMEM_BYTES));
        configured = true;
    }
    if (NB == 256) {
        if (three) (void)crnn_launch(xty_gemm_tc_kernel<256, false>, grid, DW_THREADS, DwCfg<256>::SMEM_BYTES, st, a);
        else (void)crnn_launch(xty_gemm_tc_kernel<256, true>, grid, DW_THREADS, DwCfg<256>::SMEM_BYTES, st, a);
    } else {
        if (three) (void)crnn_launch(xty_gemm_tc_kernel<128, false>, grid, DW_THREADS, DwCfg<128>::SMEM_BYTES, st, a);
        else (void)crnn_launch(xty_gemm_tc_kernel<128, true>, grid, DW_THREADS, DwCfg<128>::SMEM_BYTES, st, a);
    }
    LAUNCH_CHECK();
    return CRNN_OK;
}
