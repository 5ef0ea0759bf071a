// sddmm.cu -- sampled dense-dense product on the CSR pattern, sm_100a:
//     out_val[e] = < a[row(e), :], x[col[e], :] >      (optionally / max(deg(row(e)), 1))
//
// This is d(loss)/d(value[e]) of the sum (resp. mean) SpMM with a = grad_out.  The reference
// leaves that gradient undefined (/root/reference/csrc/fusedmm.cpp:268-272, :349-353) although
// its message header already has the ROP_DOT stage for it (csrc/fusedMM.h:34); SURVEY.md
// section 8f ranks it as the next op after the hot path.  Same access pattern and roofline as
// the forward SpMM: one K-row gather of x per stored entry.
//
// One warp owns one plan segment (same plan as the forward).  The a-row lives in registers;
// x rows are gathered exactly like in spmm_seg_kernel (coalesced index chunk, shuffles,
// 16-byte loads, U gathers in flight).  Each lane accumulates, for every entry of the chunk
// that its lane group handles, the partial dot product over its own K slice; a halving
// butterfly (G-1 shuffles per group for G entries, i.e. ~1 shuffle per entry) then leaves
// lane (group g, lane lg) with the full dot product of entry lg*NG+g, which is stored
// coalesced.  Deterministic: fixed reduction tree, no atomics.
#include "common.cuh"

namespace isplib {

struct SddmmParams {
    const int32_t* __restrict__ rowptr;
    const int32_t* __restrict__ col;
    const float* __restrict__ a;      // [m, k]
    const float* __restrict__ x;      // [n, k]
    float* __restrict__ out;          // [nnz]
    const int4* __restrict__ item_desc;   // {row, eb, ee, .} per work item (the forward's plan)
    long long lda, ldx;
    int k, num_items, seg_len, mean_scale;
    int accum;   // 1: add to out (second and later K tiles), 0: overwrite
};

template <int VEC>
__device__ __forceinline__ void ld_vec(const float* __restrict__ p, float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = __ldg(p);
    }
}

// G lanes per entry, 32/G entries per step; each lane covers LPL vectors of VEC floats per pass
template <int VEC, int G, int LPL>
__global__ void __launch_bounds__(128)
sddmm_seg_kernel(const __grid_constant__ SddmmParams p) {
    constexpr int NG = 32 / G;
    constexpr int P = (G < 16) ? G : 16; // entries per lane group per sub-chunk (partials held per lane)
    constexpr int SUB = 32 / (P * NG);   // sub-chunks per 32-entry index chunk
    constexpr int U = (P >= 4) ? 4 : P;  // gathers in flight
    constexpr int PASS_W = G * LPL * VEC;
    constexpr unsigned FULL = 0xffffffffu;

    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= p.num_items) return;
    const int4 desc = __ldg(p.item_desc + item);
    const int row = desc.x, eb = desc.y, ee = desc.z;
    if (eb >= ee) return;
    const int g = lane / G, lg = lane % G;
    float inv = 1.f;
    if (p.mean_scale) inv = __fdiv_rn(1.f, (float)max(__ldg(p.rowptr + row + 1) - __ldg(p.rowptr + row), 1));
    const unsigned ldxb = (unsigned)p.ldx * 4u;
    const int keff = (VEC == 4) ? ((p.k + 3) & ~3) : p.k;
    const float* arow = p.a + (size_t)row * (size_t)p.lda;
    const bool single_pass = keff <= PASS_W;       // the whole a-row fits the lanes once: keep it in registers

    float av[LPL][VEC];
    int ko[LPL];
    bool ok[LPL];
    auto load_a = [&](int kp) {
#pragma unroll
        for (int j = 0; j < LPL; ++j) {
            ko[j] = kp + (lg + j * G) * VEC;
            ok[j] = ko[j] < keff;
#pragma unroll
            for (int q = 0; q < VEC; ++q) av[j][q] = 0.f;
            if (ok[j]) {
                if (VEC == 1 || ko[j] + VEC <= p.k) ld_vec<VEC>(arow + ko[j], av[j]);
                else {
#pragma unroll
                    for (int q = 0; q < VEC; ++q) if (ko[j] + q < p.k) av[j][q] = __ldg(arow + ko[j] + q);
                }
            }
            if (!ok[j]) ko[j] = 0;   // out-of-range lanes re-read the row's first vector (times 0)
        }
    };
    if (single_pass) load_a(0);

    unsigned c_next = (eb + lane < ee) ? (unsigned)__ldcs(p.col + eb + lane) : 0u;
    for (int e0 = eb; e0 < ee; e0 += 32) {
        const int cnt = min(32, ee - e0);
        const unsigned c = c_next;
        if (e0 + 32 + lane < ee) c_next = (unsigned)__ldcs(p.col + e0 + 32 + lane);   // next chunk's indices
#pragma unroll
        for (int sc = 0; sc < SUB; ++sc) {
            const int sbase = sc * P * NG;
            if (sbase >= cnt) break;                       // warp-uniform
            float part[P];
#pragma unroll
            for (int v = 0; v < P; ++v) part[v] = 0.f;

            for (int kp = 0; kp < keff; kp += PASS_W) {
                if (!single_pass) load_a(kp);
                if (cnt - sbase >= P * NG) {
#pragma unroll
                    for (int v0 = 0; v0 < P; v0 += U) {
                        float xv[U][LPL][VEC];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const unsigned cc = __shfl_sync(FULL, c, sbase + (v0 + u) * NG + g);
                            const float* xr = reinterpret_cast<const float*>(reinterpret_cast<const char*>(p.x) + (unsigned long long)cc * ldxb);
#pragma unroll
                            for (int j = 0; j < LPL; ++j) ld_vec<VEC>(xr + ko[j], xv[u][j]);
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u)
#pragma unroll
                            for (int j = 0; j < LPL; ++j) {
                                // x's own padding (K % 4 != 0) may hold anything: mask it, do not multiply it
                                if (VEC == 4 && ok[j] && ko[j] + VEC > p.k) {
#pragma unroll
                                    for (int q = 0; q < VEC; ++q) if (ko[j] + q >= p.k) xv[u][j][q] = 0.f;
                                }
#pragma unroll
                                for (int q = 0; q < VEC; ++q) part[v0 + u] = fmaf(av[j][q], xv[u][j][q], part[v0 + u]);
                            }
                    }
                } else {
#pragma unroll
                    for (int v0 = 0; v0 < P; v0 += U) {
                        float xv[U][LPL][VEC];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int idx = sbase + (v0 + u) * NG + g;
                            const unsigned cc = __shfl_sync(FULL, c, idx & 31);
                            const float* xr = reinterpret_cast<const float*>(reinterpret_cast<const char*>(p.x) + (unsigned long long)cc * ldxb);
#pragma unroll
                            for (int j = 0; j < LPL; ++j) {
                                if (idx < cnt) {
                                    ld_vec<VEC>(xr + ko[j], xv[u][j]);
                                    if (VEC == 4 && ok[j] && ko[j] + VEC > p.k) {
#pragma unroll
                                        for (int q = 0; q < VEC; ++q) if (ko[j] + q >= p.k) xv[u][j][q] = 0.f;
                                    }
                                } else {
#pragma unroll
                                    for (int q = 0; q < VEC; ++q) xv[u][j][q] = 0.f;
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u)
#pragma unroll
                            for (int j = 0; j < LPL; ++j)
#pragma unroll
                                for (int q = 0; q < VEC; ++q) part[v0 + u] = fmaf(av[j][q], xv[u][j][q], part[v0 + u]);
                    }
                }
            }

            // halving butterfly inside each lane group: P values -> 1 value per lane over the low
            // lane bits, then plain adds over the remaining lane bits (G > P); lanes lg < P end
            // with the total of value index lg
#pragma unroll
            for (int o = P / 2; o >= 1; o >>= 1) {
                const bool hi = (lg & o) != 0;
#pragma unroll
                for (int h = 0; h < o; ++h) {
                    const float keep = hi ? part[h + o] : part[h];
                    const float send = hi ? part[h] : part[h + o];
                    part[h] = keep + __shfl_xor_sync(FULL, send, o);
                }
            }
#pragma unroll
            for (int o = P; o < G; o <<= 1) part[0] += __shfl_xor_sync(FULL, part[0], o);
            const int idx = sbase + lg * NG + g;
            if (lg < P && idx < cnt) {
                const float r = part[0] * inv;
                p.out[e0 + idx] = p.accum ? p.out[e0 + idx] + r : r;
            }
        }
    }
}


// ------------------------------------------------------------------------------------
// lean 256-bit variant (same recipe as spmm_lean_kernel): 32-byte gathers, U = 4 in flight, the
// step loop rolled so that only 4 x 8 staging registers are live.  G lanes cover one tile of up
// to G x 8 floats of an x row, 32/G entries per step.  Every 4 steps the 4 partial dot products
// of a lane are folded over the lane group right away (3 butterfly shuffles + log2(G/4) adds),
// so nothing but one result register is carried across iterations; lane lg ends up holding the
// dot product of step lg and the chunk is stored with one coalesced store.
// RAGGED: the tile is narrower than G x 8 floats or ends inside a vector (K = 200, 47 ...):
// surplus lanes and the columns past k are masked to zero (x's padding may hold anything).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void ld_vec8(const float* __restrict__ p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}

template <int G, bool RAGGED>
__global__ void __launch_bounds__(128, 8)
sddmm_lean256_kernel(const __grid_constant__ SddmmParams p) {
    constexpr int VEC = 8, U = 4;
    constexpr int NG = 32 / G;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= p.num_items) return;
    const int4 desc = __ldg(p.item_desc + item);
    const int row = desc.x, eb = desc.y, ee = desc.z;
    if (eb >= ee) return;
    const int g = lane / G, lg = lane % G;
    float inv = 1.f;
    if (p.mean_scale) inv = __fdiv_rn(1.f, (float)max(__ldg(p.rowptr + row + 1) - __ldg(p.rowptr + row), 1));
    const unsigned ldxb = (unsigned)p.ldx * 4u;
    const int nvalid = RAGGED ? max(0, min(VEC, p.k - lg * VEC)) : VEC;   // columns of this lane's vector inside [0, k)
    const int koff = (nvalid > 0) ? lg * VEC : 0;                          // surplus lanes park on the first vector
    const char* const xlane = reinterpret_cast<const char*>(p.x + koff);

    float av[VEC];
    {
        const float* arow = p.a + (size_t)row * (size_t)p.lda + koff;
#pragma unroll
        for (int q = 0; q < VEC; ++q) av[q] = (q < nvalid) ? __ldg(arow + q) : 0.f;
    }

    // lanes past the segment end gather row 0 (valid memory, nnz > 0 implies n > 0); their
    // results are never stored.  The next chunk's indices are fetched one chunk ahead.
    unsigned c_next = (eb + lane < ee) ? (unsigned)__ldcs(p.col + eb + lane) : 0u;
    for (int e0 = eb; e0 < ee; e0 += 32) {
        const int cnt = min(32, ee - e0);
        const unsigned c = c_next;
        c_next = (e0 + 32 + lane < ee) ? (unsigned)__ldcs(p.col + e0 + 32 + lane) : 0u;
        float res = 0.f;
#pragma unroll 1
        for (int t = 0; t < G; t += U) {
            if (t * NG >= cnt) break;                      // warp-uniform
            float xv[U][VEC];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned cc = __shfl_sync(FULL, c, ((t + u) * NG + g) & 31);
                ld_vec8(reinterpret_cast<const float*>(xlane + (unsigned long long)cc * ldxb), xv[u]);
            }
            float part[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                part[u] = 0.f;
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    if (RAGGED && q >= nvalid) xv[u][q] = 0.f;
                    part[u] = fmaf(av[q], xv[u][q], part[u]);
                }
            }
            // 4 values -> 1 over lane bits 1 and 0 of the group, then plain adds over the rest
            {
                const bool hi2 = (lg & 2) != 0;
                const float k0 = hi2 ? part[2] : part[0], s0 = hi2 ? part[0] : part[2];
                const float k1 = hi2 ? part[3] : part[1], s1 = hi2 ? part[1] : part[3];
                const float a0 = k0 + __shfl_xor_sync(FULL, s0, 2);   // step (lg&2) + 0
                const float a1 = k1 + __shfl_xor_sync(FULL, s1, 2);   // step (lg&2) + 1
                const bool hi1 = (lg & 1) != 0;
                const float kk = hi1 ? a1 : a0, ss = hi1 ? a0 : a1;
                float tot = kk + __shfl_xor_sync(FULL, ss, 1);        // step lg & 3
#pragma unroll
                for (int o = 4; o < G; o <<= 1) tot += __shfl_xor_sync(FULL, tot, o);
                if ((lg >> 2) == (t >> 2)) res = tot;                 // lane lg keeps step lg
            }
        }
        const int idx = lg * NG + g;
        if (idx < cnt) {
            const float r = res * inv;
            p.out[e0 + idx] = p.accum ? p.out[e0 + idx] + r : r;
        }
    }
}

}  // namespace isplib

using namespace isplib;

extern "C" int isplib_b200_sddmm_csr(int64_t m, int64_t n, int64_t k, int64_t nnz,
                                     const int32_t* rowptr, const int32_t* col,
                                     const float* a, int64_t lda, const float* x, int64_t ldx,
                                     int mean_scale, float* out_val,
                                     const isplib_b200_plan_info* info, const void* plan_dev,
                                     isplib_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (m < 0 || n < 0 || k < 0 || nnz < 0 || m >= INT32_MAX - 1 || nnz >= INT32_MAX || k >= INT32_MAX) return ISPLIB_INVALID_ARG;
    if (nnz == 0 || m == 0) return ISPLIB_SUCCESS;
    if (!rowptr || !col || !out_val || !info || !plan_dev) return ISPLIB_INVALID_ARG;
    if (info->m != m || info->nnz != nnz) return ISPLIB_FAIL;
    if (k == 0) { ISPLIB_CUDA_TRY(cudaMemsetAsync(out_val, 0, (size_t)nnz * 4, stream)); return ISPLIB_SUCCESS; }
    if (!a || !x || lda < k || ldx < k) return ISPLIB_INVALID_ARG;

    const PlanLayout L = plan_layout(m, nnz, info->seg_len);
    const char* base = (const char*)plan_dev;
    SddmmParams p;
    p.rowptr = rowptr; p.col = col; p.a = a; p.x = x; p.out = out_val;
    p.item_desc = (const int4*)(base + L.off_item_desc);
    p.lda = lda; p.ldx = ldx; p.k = (int)k; p.num_items = (int)info->num_items;
    p.seg_len = info->seg_len; p.mean_scale = mean_scale;

    const bool vec4 = (ldx % 4 == 0) && (ldx >= ((k + 3) & ~(int64_t)3)) && (lda % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(a) & 15u) == 0);
    // K tiling as in the forward: when x does not fit L2 but an [n, 64] slab does, sweep the
    // slabs one launch after the other; the dot products accumulate into out_val in place
    // (launch order = tile order, so the result is deterministic).
    int64_t kt = k;
    if (vec4 && k > 64 && k % 4 == 0 && (double)n * (double)k * 4.0 > 96.0 * 1024 * 1024 &&
        (double)n * 64.0 * 4.0 <= 64.0 * 1024 * 1024)
        kt = 64;
    // 32-byte gathers need 32-byte aligned x rows that hold whole vectors (the op layer pads odd
    // K to 8); tiles of at most 256 floats per launch
    const bool lean = (ldx % 8 == 0) && (ldx >= ((k + 7) & ~(int64_t)7)) && ((reinterpret_cast<uintptr_t>(x) & 31u) == 0) &&
                      k >= 32 && !getenv("ISPLIB_B200_SDDMM_SEG");
    if (lean && kt > 256) kt = 256;
    const int warps = 4;
    const dim3 grid((unsigned)((p.num_items + warps - 1) / warps)), block(warps * 32);
    for (int64_t k0 = 0; k0 < k; k0 += kt) {
        const int64_t kw = (k - k0 < kt) ? (k - k0) : kt;
        p.a = a + k0;
        p.x = x + k0;
        p.k = (int)kw;
        p.accum = k0 > 0 ? 1 : 0;
        if (lean) {
            const bool ragged = (kw != 32 && kw != 64 && kw != 128 && kw != 256);
            if (kw <= 32) { if (ragged) sddmm_lean256_kernel<4, true><<<grid, block, 0, stream>>>(p); else sddmm_lean256_kernel<4, false><<<grid, block, 0, stream>>>(p); }
            else if (kw <= 64) { if (ragged) sddmm_lean256_kernel<8, true><<<grid, block, 0, stream>>>(p); else sddmm_lean256_kernel<8, false><<<grid, block, 0, stream>>>(p); }
            else if (kw <= 128) { if (ragged) sddmm_lean256_kernel<16, true><<<grid, block, 0, stream>>>(p); else sddmm_lean256_kernel<16, false><<<grid, block, 0, stream>>>(p); }
            else { if (ragged) sddmm_lean256_kernel<32, true><<<grid, block, 0, stream>>>(p); else sddmm_lean256_kernel<32, false><<<grid, block, 0, stream>>>(p); }
        } else if (vec4) {
            const int64_t tv = ((kw + 3) & ~(int64_t)3) / 4;
            if (tv <= 8) sddmm_seg_kernel<4, 8, 1><<<grid, block, 0, stream>>>(p);
            else if (tv <= 16) sddmm_seg_kernel<4, 16, 1><<<grid, block, 0, stream>>>(p);
            else if (tv <= 32) sddmm_seg_kernel<4, 32, 1><<<grid, block, 0, stream>>>(p);
            else sddmm_seg_kernel<4, 32, 2><<<grid, block, 0, stream>>>(p);   // 256 floats per pass, loops for wider K
        } else {
            if (kw <= 32) sddmm_seg_kernel<1, 32, 1><<<grid, block, 0, stream>>>(p);
            else sddmm_seg_kernel<1, 32, 2><<<grid, block, 0, stream>>>(p);
        }
        ISPLIB_LAUNCH_CHECK();
    }
    return ISPLIB_SUCCESS;
}
