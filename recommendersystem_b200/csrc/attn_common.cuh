// Shared pieces of the attention kernels (attn_sm100.cu: one head per CTA;
// attn_pair_sm100.cu: two heads of one GQA group per CTA, one warpgroup each).
#pragma once
#include "common.cuh"

namespace graft {

constexpr float kScale = 0.125f;                       // 1/sqrt(64)
constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;

struct AttnDev {
  int H, G, rows, row_len;
  long long T;
  const int32_t* tm; const int32_t* seg_start; const int32_t* seg_end;
  int k_col, v_col;
  __nv_bfloat16* out; float* lse2;          // fwd outputs
  const float* delta;                        // bwd
  __nv_bfloat16* dqkv; long long ld;         // bwd output (layout of qkv)
  int mode;                                  // 0 softmax, 1 point-wise SiLU (pair kernels only)
  const float* rope_cos; const float* rope_sin;  // bwd (optional): [max_pos, 32] tables; dq / dk are stored un-rotated
  unsigned int* flags;                       // the library's status word (graft_status_flags)
  const uint32_t* zero_words; const uint32_t* one_words;  // [T/32] bitmaps of tm == 0 / tm == 1 (graft_token_mask_words)
};

// byte offset of (row, 16-byte chunk) inside a [rows x 128 B] SW128 K-major tile
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_ld_32x32_f(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld_32x32(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

// two 32-column loads in flight, one wait: the second load's latency hides behind the first's
__device__ __forceinline__ void tmem_ld_32x32_f2(uint32_t taddr_a, float (&a)[32], uint32_t taddr_b, float (&b)[32]) {
  uint32_t ra[32], rb[32];
  tmem_ld_32x32(taddr_a, ra);
  tmem_ld_32x32(taddr_b, rb);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) { a[j] = __uint_as_float(ra[j]); b[j] = __uint_as_float(rb[j]); }
}

// bits [lo,hi) of the tile's columns that fall into 32-column word w
__device__ __forceinline__ uint32_t range_word(int lo, int hi, int w) {
  const int a = min(max(lo - 32 * w, 0), 32), b = min(max(hi - 32 * w, 0), 32);
  const uint32_t mb = (b >= 32) ? 0xffffffffu : ((1u << b) - 1u);
  const uint32_t ma = (a >= 32) ? 0xffffffffu : ((1u << a) - 1u);
  return mb & ~ma;
}
// columns of a word (inside `range`) whose mask id equals `id`; only mask-id != 0 columns are candidates
__device__ __forceinline__ uint32_t equal_id_word(uint32_t range, uint32_t zero_word, const int32_t* tm_cols, int id) {
  uint32_t nz = range & ~zero_word, e = 0;
  while (nz) {
    const int c = __ffs(nz) - 1;
    if (tm_cols[c] == id) e |= 1u << c;
    nz &= nz - 1;
  }
  return e;
}

// barrier among the 128 threads of one warpgroup (ids 1,2; 0 is __syncthreads)
__device__ __forceinline__ void wg_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

__device__ __forceinline__ void store_row_bf16_32(__nv_bfloat16* dst, const float (&v)[32], float mul) {
#pragma unroll
  for (int q8 = 0; q8 < 4; ++q8)
    reinterpret_cast<uint4*>(dst)[q8] =
        make_uint4(pack_bf16(v[8 * q8] * mul, v[8 * q8 + 1] * mul), pack_bf16(v[8 * q8 + 2] * mul, v[8 * q8 + 3] * mul),
                   pack_bf16(v[8 * q8 + 4] * mul, v[8 * q8 + 5] * mul), pack_bf16(v[8 * q8 + 6] * mul, v[8 * q8 + 7] * mul));
}

// backward of RoPE (transformer.model.py:182-190) on 32 consecutive dims (16 pairs: dims 32*half ...) of one head
// row held in registers: the transpose rotation x0' = x0 c + x1 s, x1' = -x0 s + x1 c at the token's position
__device__ __forceinline__ void unrope_row_32(float (&v)[32], const AttnDev& p, int pos, int half) {
  const float4* c4 = reinterpret_cast<const float4*>(p.rope_cos + (long long)pos * 32 + half * 16);
  const float4* s4 = reinterpret_cast<const float4*>(p.rope_sin + (long long)pos * 32 + half * 16);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 c = __ldg(c4 + q), s = __ldg(s4 + q);
    const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = 8 * q + 2 * e;
      const float x0 = v[j], x1 = v[j + 1];
      v[j] = x0 * cc[e] + x1 * ss[e];
      v[j + 1] = x1 * cc[e] - x0 * ss[e];
    }
  }
}

// implemented in attn_pair_sm100.cu; used when H/G is even
int launch_attn_fwd_pair(const CUtensorMap& tmQKV128, const AttnDev& p, cudaStream_t st);
int launch_attn_bwd_pair(const CUtensorMap& tq128, const CUtensorMap& tq64, const CUtensorMap& td128, const CUtensorMap& td64,
                         const AttnDev& p, cudaStream_t st);

}  // namespace graft
