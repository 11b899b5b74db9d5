// b200sd -- flash-style attention for sm_100a: S = Q K^T and O += P V on tcgen05 tensor cores
// (accumulators in TMEM), Q/K/V tiles staged by TMA (SWIZZLE_128B), online softmax (exp2 domain)
// in registers, P handed back to the tensor core through shared memory.
//
// Replaces attention.original / split_einsum / split_einsum_v2 of the reference
// (python_coreml_stable_diffusion/attention.py:24-168, dispatched by Einsum unet.py:45-59): the
// three variants are one function, softmax(q^T k / sqrt(d) + mask) v per (batch, head); the
// (B, heads, Sq, Sk) score tensor the reference materialises (671 MB at S=4096) never leaves the SM.
//
// CTA = 128 queries x 1 head; warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 softmax (one query
// row per thread).  The pipeline runs on 64-key halves: S is double-buffered in TMEM (2 x 64 columns) and P
// in shared memory (2 x 16 KiB), so S_{h+1} = Q K^T and O += P_{h-1} V execute on the tensor core while the
// softmax warps exponentiate half h.  256 TMEM columns (S 128 + O 64) so two CTAs share an SM.
#include "common.cuh"
#include "../../include/b200sd.h"

namespace b200sd {

extern void count_launch(int n);

static constexpr int kQ = 128;   // queries per CTA
static constexpr int kKV = 128;  // keys per K/V tile (one TMA load)
static constexpr int kHalf = 64;  // keys per pipeline step: S is double-buffered by halves of a tile
static constexpr int kD = 64;    // head dim
static constexpr int kAttnThreads = 192;
static constexpr int kTileBytes = 128 * 64 * 2;  // 16 KiB: one [128 x 64] fp16 tile
static constexpr int kKvStages = 2;

struct __align__(64) AttnParams {
    CUtensorMap tmQ, tmK, tmV;
    __half* out;
    const float* mask;  // [batch, sk] additive or null
    int sq, sk, ldo;
    int causal;  // 1: key j is visible to query i only if j <= i (CLIP text encoder)
    float scale_log2;  // scale * log2(e)
};

// smem layout (1024-aligned): Q | P (2 x 16K, K-chunks of 64 keys) | K[2] | V[2] | barriers
static constexpr int kSmemQ = 0;
static constexpr int kSmemP = kTileBytes;
static constexpr int kSmemK = kSmemP + 2 * kTileBytes;
static constexpr int kSmemV = kSmemK + kKvStages * kTileBytes;
static constexpr int kSmemBar = kSmemV + kKvStages * kTileBytes;
static constexpr int kAttnSmemBytes = kSmemBar + 128;

__global__ void __launch_bounds__(kAttnThreads, 2) attention_kernel(const __grid_constant__ AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
    uint64_t* q_full = bars + 0;
    uint64_t* kv_full = bars + 1;   // [2]
    uint64_t* kv_empty = bars + 3;  // [2]
    uint64_t* s_full = bars + 5;    // [2]  per 64-key half of S
    uint64_t* s_empty = bars + 7;   // [2]
    uint64_t* p_full = bars + 9;    // [2]  per 64-key chunk of P
    uint64_t* o_full = bars + 11;   // [2]  P V of the chunk retired (P chunk reusable; O quiescent up to here)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kQ;
    const int head = blockIdx.y;
    const int batch = blockIdx.z;
    // the pipeline runs on 64-key halves of the 128-key tiles; under a causal mask the halves entirely above this
    // query tile's diagonal are never visited
    const int sk_eff = p.causal ? min(p.sk, q0 + kQ) : p.sk;
    const int n_kv = (sk_eff + kKV - 1) / kKV;
    const int n_half = (sk_eff + kHalf - 1) / kHalf;

    if (threadIdx.x == 0) {
        // SWIZZLE_128B tiles need a 1024-byte aligned base; a padded buffer would cost the second CTA per SM, so a
        // misaligned launch (never observed: the kernel has no static shared memory) fails loudly instead
        if ((smem_u32(smem) & 1023u) != 0) __trap();
        prefetch_tmap(&p.tmQ);
        prefetch_tmap(&p.tmK);
        prefetch_tmap(&p.tmV);
        mbar_init(q_full, 1);
        for (int s = 0; s < kKvStages; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&s_full[b], 1);
            mbar_init(&s_empty[b], 128);
            mbar_init(&p_full[b], 128);
            mbar_init(&o_full[b], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();  // PDL: the prologue above overlapped the previous kernel's tail
    const uint32_t tmem_s = tmem_base;        // 2 x 64 columns (double-buffered S halves)
    const uint32_t tmem_o = tmem_base + 128;  // 64 columns

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(q_full, kTileBytes);
            tma_load_3d(smem + kSmemQ, &p.tmQ, q_full, head * kD, q0, batch, kEvictFirst);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j % kKvStages;
                const uint32_t ph = (j / kKvStages) & 1;
                mbar_wait(&kv_empty[st], ph ^ 1);
                mbar_expect_tx(&kv_full[st], 2 * kTileBytes);
                tma_load_3d(smem + kSmemK + st * kTileBytes, &p.tmK, &kv_full[st], head * kD, j * kKV, batch,
                            kEvictLast);
                tma_load_3d(smem + kSmemV + st * kTileBytes, &p.tmV, &kv_full[st], head * kD, j * kKV, batch,
                            kEvictLast);
            }
        }
    } else if (warp == 1) {
        // S_h = Q K_h^T : A = Q (K-major), B = 64 rows of the K tile (K-major), M=128 N=64 K=64
        const uint32_t idesc_s = make_idesc_f16(128, kHalf, 0, 0);
        // O += P_h V_h : A = P chunk (K-major, 64 keys), B = 64 rows of the V tile [keys][d] = MN-major, N=64 K=64
        const uint32_t idesc_o = make_idesc_f16(128, kD, 0, 1);
        const uint32_t q_addr = smem_u32(smem + kSmemQ);
        const uint32_t p_addr = smem_u32(smem + kSmemP);
        auto issue_s = [&](int h) {
            const int j = h >> 1, b = h & 1;
            const int st = j % kKvStages;
            mbar_wait(&kv_full[st], (j / kKvStages) & 1);
            if (h >= 2) mbar_wait(&s_empty[b], ((h - 2) >> 1) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint64_t adesc = make_smem_desc_sw128(q_addr, 1024, 0);
                const uint64_t bdesc =
                    make_smem_desc_sw128(smem_u32(smem + kSmemK + st * kTileBytes) + b * (kHalf * 128), 1024, 0);
#pragma unroll
                for (int k = 0; k < kD / 16; ++k)
                    umma_f16_ss(tmem_s + b * kHalf, adesc + 2 * k, bdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
                umma_commit(&s_full[b]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        issue_s(0);
        if (n_half > 1) issue_s(1);
        for (int h = 0; h < n_half; ++h) {
            const int j = h >> 1, b = h & 1;
            const int st = j % kKvStages;
            mbar_wait(&p_full[b], (h >> 1) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t v_addr = smem_u32(smem + kSmemV + st * kTileBytes) + b * (kHalf * 128);
#pragma unroll
                for (int k = 0; k < kHalf / 16; ++k) {
                    // A: +32 B per 16 keys inside the chunk's 128 B rows;  B: 16 keys = 16 rows of 128 B = 2048 B
                    const uint64_t adesc = make_smem_desc_sw128(p_addr + b * kTileBytes, 1024, 0) + 2 * k;
                    const uint64_t bdesc = make_smem_desc_sw128(v_addr + k * 2048, 1024, kKV * 128);
                    umma_f16_ss(tmem_o, adesc, bdesc, idesc_o, (h > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&o_full[b]);
                if (b == 1 || h == n_half - 1) umma_commit(&kv_empty[st]);  // last half of this K/V tile
            }
            __syncwarp();
            if (h + 2 < n_half) issue_s(h + 2);
        }
    } else {
        // ---------------- softmax warps: one query row per thread ----------------
        // O accumulates in TMEM across all KV steps (PV MMAs with accumulate=1).  The exp2 reference m_ref
        // is refreshed lazily: a full, unmasked half tile is exponentiated optimistically against the current
        // m_ref in ONE pass over S (row maximum tracked on the side); only if some row's maximum exceeds
        // m_ref by more than kTau (log2 domain; P <= 2^kTau stays inside fp16) does the warp wait for the
        // outstanding P V MMAs, rescale its 32 rows of O in TMEM and redo the half -- rare after the first tiles.
        // S is double-buffered by halves, so the tensor core computes S_{h+1} / S_{h+2} and P_{h-1} V while
        // this half is in the exponentials.
        constexpr float kTau = 8.0f;
        const int lane_group = warp & 3;
        const int row = lane_group * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(lane_group * 32) << 16;
        const uint32_t o_addr = tmem_o + lane_addr;
        const float sl2 = p.scale_log2;
        float m_ref = -INFINITY, l_run = 0.f;
        const float* mask_row = p.mask ? p.mask + static_cast<size_t>(batch) * p.sk : nullptr;
        uint8_t* p_row = smem + kSmemP + row * 128;
        const int sw = row & 7;

        // lean row maximum of a full unmasked half (raw scores)
        auto row_max_lean = [&](uint32_t s_addr) {
            float m0 = -INFINITY, m1 = -INFINITY;
            uint32_t va[32], vb[32];
            tmem_ld32(s_addr, va);
            tmem_ld32(s_addr + 32, vb);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                m0 = fmaxf(m0, fmaxf(__uint_as_float(va[i]), __uint_as_float(vb[i])));
                m1 = fmaxf(m1, fmaxf(__uint_as_float(va[i + 1]), __uint_as_float(vb[i + 1])));
            }
            return fmaxf(m0, m1);
        };
        // lean probabilities of a full unmasked half against reference `ref`: writes the fp16 P chunk (one
        // 128-byte swizzled row per query: eight 16-byte pieces), returns the row sum and (tmax) the raw maximum
        auto probs_lean = [&](uint32_t s_addr, uint8_t* dst, float ref, float& tmax) {
            const float neg_m = -ref;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            float m0 = -INFINITY, m1 = -INFINITY;
            uint32_t va[32], vb[32];
            tmem_ld32(s_addr, va);
            tmem_ld32(s_addr + 32, vb);
            tmem_ld_wait();
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const float a0 = __uint_as_float(va[i]), a1 = __uint_as_float(va[i + 1]);
                const float b0 = __uint_as_float(vb[i]), b1 = __uint_as_float(vb[i + 1]);
                m0 = fmaxf(m0, fmaxf(a0, b0));
                m1 = fmaxf(m1, fmaxf(a1, b1));
                const float p0 = ex2_approx(fmaf(a0, sl2, neg_m));
                const float p1 = ex2_approx(fmaf(a1, sl2, neg_m));
                const float p2 = ex2_approx(fmaf(b0, sl2, neg_m));
                const float p3 = ex2_approx(fmaf(b1, sl2, neg_m));
                s0 += p0, s1 += p1, s2 += p2, s3 += p3;
                pk[i >> 1] = pack_half2(p0, p1);
                pk[16 + (i >> 1)] = pack_half2(p2, p3);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint4 val = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                *reinterpret_cast<uint4*>(dst + ((q ^ sw) << 4)) = val;
            }
            tmax = fmaxf(m0, m1);
            return (s0 + s1) + (s2 + s3);
        };
        // rescale this warp's rows of O (TMEM) and the running sum from reference m_ref to m_new
        auto rescale_o = [&](float m_new) {
            const float factor = ex2_approx(m_ref - m_new);  // 1 for rows whose reference did not move
#pragma unroll
            for (int c = 0; c < kD; c += 32) {
                uint32_t v[32];
                tmem_ld32(o_addr + c, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * factor);
                tmem_st32(o_addr + c, v);
            }
            tmem_st_wait();
            l_run *= factor;
        };

        for (int h = 0; h < n_half; ++h) {
            const int b = h & 1;
            const int kvalid = min(kHalf, p.sk - h * kHalf);  // >= 1
            // causal: a half whose last key is <= the tile's first query is fully visible to every row
            const bool diag = p.causal && (h * kHalf + kHalf - 1 > q0);
            const int qi = q0 + row;  // this thread's query index
            const uint32_t s_addr = tmem_s + lane_addr + b * kHalf;
            uint8_t* dst = p_row + b * kTileBytes;
            mbar_wait(&s_full[b], (h >> 1) & 1);
            if (h >= 2) mbar_wait(&o_full[b], ((h - 2) >> 1) & 1);  // P V_{h-2} retired: P chunk b is free
            tc_fence_after();
            // every P V issued so far (up to half h-1) has retired: O may be rescaled in place
            auto wait_o_quiescent = [&]() {
                if (h >= 1) {
                    mbar_wait(&o_full[(h - 1) & 1], ((h - 1) >> 1) & 1);
                    tc_fence_after();
                }
            };
            float l_half;
            if (mask_row == nullptr && kvalid == kHalf && !diag) {
                if (h == 0) {
                    m_ref = row_max_lean(s_addr) * sl2;
                    float unused;
                    l_half = probs_lean(s_addr, dst, m_ref, unused);
                } else {
                    float tmax;
                    l_half = probs_lean(s_addr, dst, m_ref, tmax);
                    const float m_half = tmax * sl2;
                    if (__any_sync(0xffffffffu, m_half > m_ref + kTau)) {
                        const float m_new = fmaxf(m_ref, m_half);
                        wait_o_quiescent();
                        rescale_o(m_new);
                        m_ref = m_new;
                        l_half = probs_lean(s_addr, dst, m_ref, tmax);
                    }
                }
            } else {
                // ---- general half (additive mask and/or ragged tail): two passes with per-key predicates ----
                float m_half = -INFINITY;
#pragma unroll 1
                for (int c = 0; c < kHalf; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(s_addr + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float sc = __uint_as_float(v[i]) * sl2;
                        const bool vis = c + i < kvalid && (!diag || h * kHalf + c + i <= qi);
                        if (mask_row && vis) sc += mask_row[h * kHalf + c + i] * 1.4426950408889634f;
                        if (vis) m_half = fmaxf(m_half, sc);
                    }
                }
                if (__any_sync(0xffffffffu, m_half > m_ref + kTau)) {
                    const float m_new = fmaxf(m_ref, m_half);
                    if (h > 0) {
                        wait_o_quiescent();
                        rescale_o(m_new);
                    }
                    m_ref = m_new;
                }
                l_half = 0.f;
#pragma unroll 1
                for (int c = 0; c < kHalf; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(s_addr + c, v);
                    tmem_ld_wait();
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float s0 = __uint_as_float(v[i]) * sl2, s1 = __uint_as_float(v[i + 1]) * sl2;
                        const bool vis0 = c + i < kvalid && (!diag || h * kHalf + c + i <= qi);
                        const bool vis1 = c + i + 1 < kvalid && (!diag || h * kHalf + c + i + 1 <= qi);
                        if (mask_row) {
                            if (vis0) s0 += mask_row[h * kHalf + c + i] * 1.4426950408889634f;
                            if (vis1) s1 += mask_row[h * kHalf + c + i + 1] * 1.4426950408889634f;
                        }
                        const float p0 = vis0 ? ex2_approx(s0 - m_ref) : 0.f;
                        const float p1 = vis1 ? ex2_approx(s1 - m_ref) : 0.f;
                        l_half += p0 + p1;
                        pk[i >> 1] = pack_half2(p0, p1);
                    }
                    const int piece0 = c >> 3;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 val = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                        *reinterpret_cast<uint4*>(dst + (((piece0 + q) ^ sw) << 4)) = val;
                    }
                }
            }
            l_run += l_half;
            tc_fence_before();
            mbar_arrive(&s_empty[b]);
            fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
            mbar_arrive(&p_full[b]);
        }
        // ---- normalise and store (the last commit covers every earlier P V) ----
        mbar_wait(&o_full[(n_half - 1) & 1], ((n_half - 1) >> 1) & 1);
        tc_fence_after();
        const float inv_l = 1.0f / l_run;
        const bool store = q0 + row < p.sq;
        __half* dst = p.out + (static_cast<size_t>(batch) * p.sq + (store ? q0 + row : 0)) * p.ldo + head * kD;
#pragma unroll
        for (int c = 0; c < kD; c += 32) {
            uint32_t v[32];
            tmem_ld32(o_addr + c, v);
            tmem_ld_wait();
            if (store) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 val;
                    val.x = pack_half2(__uint_as_float(v[i]) * inv_l, __uint_as_float(v[i + 1]) * inv_l);
                    val.y = pack_half2(__uint_as_float(v[i + 2]) * inv_l, __uint_as_float(v[i + 3]) * inv_l);
                    val.z = pack_half2(__uint_as_float(v[i + 4]) * inv_l, __uint_as_float(v[i + 5]) * inv_l);
                    val.w = pack_half2(__uint_as_float(v[i + 6]) * inv_l, __uint_as_float(v[i + 7]) * inv_l);
                    *reinterpret_cast<uint4*>(dst + c + i) = val;
                }
            }
        }
    }

    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace b200sd

using namespace b200sd;

extern "C" int b200sd_attention(const void* q, const void* k, const void* v, void* out, const float* mask,
                                int32_t batch, int32_t heads, int32_t sq, int32_t sk, int32_t d, int32_t ldq,
                                int32_t ldk, int32_t ldv, int32_t ldo, float scale, int32_t impl, void* stream_) {
    if (!b200sd::launch_class_enabled(2)) return 0;  // bench.py's per-class timing graphs
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    B200SD_REQUIRE(q && k && v && out, "b200sd_attention: null pointer");
    B200SD_REQUIRE(d == kD, "b200sd_attention: head dim %d not supported by this kernel (needs 64)", d);
    B200SD_REQUIRE(impl >= 0 && (impl & 0xff) <= 2 && (impl & ~0x1ff) == 0, "b200sd_attention: unknown attention implementation %d", impl);
    B200SD_REQUIRE(batch > 0 && heads > 0 && sq > 0 && sk > 0, "b200sd_attention: bad sizes");
    B200SD_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
                   "b200sd_attention: leading dimensions must be multiples of 8");
    AttnParams p;
    memset(&p, 0, sizeof(p));
    const uint32_t es[3] = {1, 1, 1};
    const uint32_t box[3] = {kD, 128, 1};
    const void* ptrs[3] = {q, k, v};
    const int lds[3] = {ldq, ldk, ldv};
    const int seqs[3] = {sq, sk, sk};
    CUtensorMap* maps[3] = {&p.tmQ, &p.tmK, &p.tmV};
    for (int i = 0; i < 3; ++i) {
        const uint64_t dims[3] = {static_cast<uint64_t>(heads) * kD, static_cast<uint64_t>(seqs[i]),
                                  static_cast<uint64_t>(batch)};
        const uint64_t str[2] = {static_cast<uint64_t>(lds[i]) * 2, static_cast<uint64_t>(lds[i]) * 2 * seqs[i]};
        if (int rc = encode_tmap_f16(maps[i], ptrs[i], 3, dims, str, box, es)) return rc;
    }
    p.out = reinterpret_cast<__half*>(out);
    p.mask = mask;
    p.sq = sq;
    p.sk = sk;
    p.ldo = ldo;
    p.causal = (impl & 0x100) ? 1 : 0;
    p.scale_log2 = scale * 1.4426950408889634f;
    static bool attr_set = false;
    if (!attr_set) {
        B200SD_CHECK_CUDA(
            cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
        attr_set = true;
    }
    dim3 grid((sq + kQ - 1) / kQ, heads, batch);
    B200SD_CHECK_CUDA(launch_kernel(attention_kernel, dim3(grid), dim3(kAttnThreads), kAttnSmemBytes, stream, p));
    B200SD_CHECK_CUDA(cudaGetLastError());
    count_launch(1);
    return 0;
}
