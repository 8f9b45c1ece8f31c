"""
Derives scratch/ffae_infer_tc_v17_layer0_128col.cu from the production kernel: layer 0 with a 128-column TMEM slot
(DESIGN section 7.1).  A_hi comes from the x box through an SS-form tf32 MMA (as scratch/ffae_infer_tc_v16_ahi_from_xbox.cu,
verified on the GPU), A_lo travels as packed BF16 against a BF16 image of W_hi.  Still two tile slots: the point of this
step is to validate numerics and the x-box life time on the GPU before the schedule grows a third slot.

    python scratch/make_v17.py && nvcc -gencode arch=compute_100a,code=sm_100a -c scratch/ffae_infer_tc_v17_layer0_128col.cu -o /tmp/v17.o -I gordo_components_b200/csrc

NOT YET RUN on a B200 (written after round 1's GPU budget was spent).  To try it: copy over csrc/ffae_infer_tc.cu, rebuild,
run `pytest -m gpu -k "tc or infer"`, then bench.
"""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(HERE, "..", "gordo_components_b200", "csrc", "ffae_infer_tc.cu")).read()


def sub(old, new, count=1):
    global src
    assert src.count(old) == count, (src.count(old), old[:80])
    src = src.replace(old, new)


sub('#include "gb_common.cuh"', '#include "../gordo_components_b200/csrc/gb_common.cuh"')
# ---- TMEM map: D 64 | bf16(A_lo) 32 | bf16(A) 32  (the later layers' FP16 pair overlays the two BF16 images)
sub("constexpr uint32_t COL_D = 0, COL_AHI = 64, COL_ALO = 128, COL_ABF = 192, SLOT_COLS = 224, COL_DX = 448, TMEM_COLS = 512;",
    "constexpr uint32_t COL_D = 0, COL_ALB = 64, COL_ABF = 96, SLOT_COLS = 128, COL_DX = 256, TMEM_COLS = 512;  // 256 columns still free: two more slots")
sub("constexpr uint32_t COL_A1 = COL_AHI, COL_A2 = COL_AHI + 32;", "constexpr uint32_t COL_A1 = COL_ALB, COL_A2 = COL_ABF;")
sub("  int whi_ofs[MAXL], wlo_ofs[MAXL], bias_ofs[MAXL];  // byte offsets into dynamic smem",
    "  int whi_ofs[MAXL], wlo_ofs[MAXL], bias_ofs[MAXL];  // byte offsets into dynamic smem\n  int whb_ofs;                                       // layer 0: BF16 image of W_hi [K/8][Np][8]")
# ---- SS-form MMA + SWIZZLE_128B descriptor of the x box
sub("__device__ __forceinline__ void mma_commit(uint32_t bar) {", '''// A and B from shared memory (layer 0's A_hi is the TMA'd x box itself: the tensor core ignores the low 13 mantissa bits of fp32 data)
__device__ __forceinline__ void mma_tf32_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\\n\\t"
      ".reg .pred p;\\n\\t"
      "setp.ne.b32 p, %4, 0;\\n\\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\\n\\t"
      "}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major SWIZZLE_128B operand (a TMA box of 128-byte rows): 8-row groups 1024 bytes apart; K steps advance the start address
__device__ __forceinline__ uint64_t make_adesc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {''')
# ---- x split: two packed BF16 images only
a = src.index("// layer 0: split NC (16) inputs into the three A operands")
b = src.index("// layers >= 1: split NC activations")
src = src[:a] + '''// layer 0: NC (16) inputs -> packed BF16 images of A_lo = A - trunc_tf32(A) and of A itself, at column `col`
// (A_hi is not stored: the tensor core reads it straight from the x box)
template <int NC>
__device__ __forceinline__ void store_a_operands(uint32_t slot_lane, int col, const float* a) {
  uint32_t lb[NC / 2], bf[NC / 2];
#pragma unroll
  for (int i = 0; i < NC / 2; ++i) {
    const float l0 = a[2 * i] - __uint_as_float(__float_as_uint(a[2 * i]) & 0xffffe000u);
    const float l1 = a[2 * i + 1] - __uint_as_float(__float_as_uint(a[2 * i + 1]) & 0xffffe000u);
    const __nv_bfloat162 pl = __floats2bfloat162_rn(l0, l1);  // low half = even k (the order the MMA expects)
    const __nv_bfloat162 pa = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]);
    lb[i] = *reinterpret_cast<const uint32_t*>(&pl);
    bf[i] = *reinterpret_cast<const uint32_t*>(&pa);
  }
  constexpr int C8 = NC / 8, R4 = (NC % 8) / 4;  // NC = 8*C8 + 4*R4
#pragma unroll
  for (int c = 0; c < C8; ++c) {
    tmem_st4(slot_lane + COL_ALB + ((col + 8 * c) >> 1), lb + 4 * c);
    tmem_st4(slot_lane + COL_ABF + ((col + 8 * c) >> 1), bf + 4 * c);
  }
  if (R4) {
    tmem_st2(slot_lane + COL_ALB + ((col + 8 * C8) >> 1), lb + 4 * C8);
    tmem_st2(slot_lane + COL_ABF + ((col + 8 * C8) >> 1), bf + 4 * C8);
  }
}

''' + src[b:]
# ---- weight staging: third image for layer 0
sub("        __nv_bfloat16* wlo = reinterpret_cast<__nv_bfloat16*>(smem + a.wlo_ofs[l]);",
    "        __nv_bfloat16* wlo = reinterpret_cast<__nv_bfloat16*>(smem + a.wlo_ofs[l]);\n        __nv_bfloat16* whb = reinterpret_cast<__nv_bfloat16*>(smem + a.whb_ofs);")
sub("                wlo[((k >> 3) * Np + n) * 8 + (k & 7)] = __float2bfloat16_rn(w - hi);",
    "                wlo[((k >> 3) * Np + n) * 8 + (k & 7)] = __float2bfloat16_rn(w - hi);\n                whb[((k >> 3) * Np + n) * 8 + (k & 7)] = __float2bfloat16_rn(hi);")
# ---- MMA issue of layer 0, and the x boxes live until those MMAs have completed
sub("            if (l == 0 && t + 2 < n_tiles) {  // A0 is in TMEM => this slot's x boxes are free",
    "            if (l == 1 && t + 2 < n_tiles) {  // layer 0's MMAs (which read the x boxes) are complete => the boxes are free")
a = src.index("              for (int ks = 0; ks < 8; ++ks)  // A_lo * W_hi (first MMA overwrites the accumulator)")
b = src.index("            } else {\n#pragma unroll\n              for (int ks = 0; ks < 4; ++ks)  // a2 * w1")
src = src[:a] + '''              for (int ks = 0; ks < 8; ++ks)  // A_hi * W_hi: A is the x box (SWIZZLE_128B, 32 columns per box, 32 bytes per K step); first MMA overwrites
                if (ks < k8) mma_tf32_ss(dcol, make_adesc_sw128(xdst + (ks >> 2) * BOX_BYTES) + (uint64_t)((ks & 3) * 2), dhi + (uint64_t)(ks * dstep), id32, ks > 0);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)  // bf16(A_lo) * bf16(W_hi)
                if (ks < k16) mma_bf16_ts(dcol, tb + COL_ALB + ks * 8, dhb + (uint64_t)(ks * dstep), id16, 1);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)  // bf16(A) * bf16(W_lo)
                if (ks < k16) mma_bf16_ts(dcol, tb + COL_ABF + ks * 8, dlo + (uint64_t)(ks * dstep), id16, 1);
''' + src[b:]
sub("          const uint64_t dhi = make_bdesc(sbase + a.whi_ofs[l], lbo, 128), dlo = make_bdesc(sbase + a.wlo_ofs[l], lbo, 128);",
    "          const uint64_t dhi = make_bdesc(sbase + a.whi_ofs[l], lbo, 128), dlo = make_bdesc(sbase + a.wlo_ofs[l], lbo, 128), dhb = make_bdesc(sbase + a.whb_ofs, lbo, 128);")
# the x boxes are re-armed after layer 0 (at l == 1): a one-layer stack would never prefetch
sub("  if (L > MAXL || net->dims[0] != net->dims[L] ||", "  if (L < 2 || L > MAXL || net->dims[0] != net->dims[L] ||")
# ---- host: room for the third image
sub("    a.wlo_ofs[l] = ofs;\n    ofs += a.k16[l] * 16 * a.Np[l] * 2;\n",
    "    a.wlo_ofs[l] = ofs;\n    ofs += a.k16[l] * 16 * a.Np[l] * 2;\n    if (l == 0) {\n      a.whb_ofs = ofs;\n      ofs += a.k16[l] * 16 * a.Np[l] * 2;\n    }\n")
open(os.path.join(HERE, "ffae_infer_tc_v17_layer0_128col.cu"), "w").write(src)
print("written")
