"""
Derives scratch/ffae_infer_tc_v18_three_slots.cu from the v17 prototype (scratch/make_v17.py): THREE 128-row tiles in flight.

TMEM: three 128-column slots (accumulator 64 + the two packed 16-bit operand images 32 + 32) + two spare accumulators (slots 1
and 2 send their output layer there) = 512 columns.  Shared memory: 88 KB weight images + 96 KB x boxes (3 slots x 2) + 32 KB
transpose staging = 221 KB; the per-warp y box of the two-slot kernel no longer fits, so the y rows of a triple's second and third
tile are requested into registers as soon as the previous tile's stores have consumed them.  19 warps: 8 layer, 8 output, 3 control.

NOT YET RUN on a B200 (written after round 1's GPU budget was spent).  Expected from the cycle trace of the two-slot kernel
(DESIGN section 4.1): the issue/commit/epilogue latency chain of a tile (~22 k cycles) is overlapped three ways instead of two,
which moves the bound to the layer warps' epilogue time (~8 k cycles per tile) -- about the HBM floor (8.5 k).
To try it: python scratch/make_v17.py && python scratch/make_v18.py, copy over csrc/ffae_infer_tc.cu (fix the include), rebuild,
run `timeout 120 pytest -m gpu -k "infer or tc"` FIRST (a barrier mistake here hangs the GPU), then bench.
"""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(HERE, "ffae_infer_tc_v17_layer0_128col.cu")).read()


def sub(old, new, count=1):
    global src
    assert src.count(old) == count, (src.count(old), old[:90])
    src = src.replace(old, new)


# ---------------------------------------------------------------- constants
sub("constexpr int NTHREADS = 576;  // warps 0-7: layer epilogues (SFU-bound); 8-15: output epilogue (LSU-bound); 16,17: control of slot 0 / 1",
    "constexpr int NSLOT = 3;       // tiles in flight\nconstexpr int NTHREADS = 608;  // warps 0-7: layer epilogues (SFU-bound); 8-15: output epilogue (LSU-bound); 16,17,18: control of slot 0 / 1 / 2")
sub("constexpr uint32_t COL_D = 0, COL_ALB = 64, COL_ABF = 96, SLOT_COLS = 128, COL_DX = 256, TMEM_COLS = 512;  // 256 columns still free: two more slots",
    "constexpr uint32_t COL_D = 0, COL_ALB = 64, COL_ABF = 96, SLOT_COLS = 128, COL_DX = 384, TMEM_COLS = 512;  // 3 slots + 2 spare accumulators (COL_DX, COL_DX + 64)")
# the debug timeline shrinks: 226 304 bytes of dynamic shared memory (BASELINE net) leave ~6 KB for static allocations
sub("constexpr int TRACE_SLOTS = 320;", "constexpr int TRACE_SLOTS = 64;  // (scratch/dbg_trace.py must be told: the buffer layout depends on it)")
sub('  GB_REQUIRE(smem <= 227 * 1024, GB_E_SMEM, "architecture needs %zu bytes of shared memory in the tcgen05 variant", smem);',
    '  GB_REQUIRE(smem + 4 * TRACE_SLOTS * 8 + 64 <= 227 * 1024, GB_E_SMEM, "architecture needs %zu bytes of shared memory in the tcgen05 variant", smem);')
# ---------------------------------------------------------------- barriers and phases
sub("  const uint32_t BX = 0, BA = 16, BD = 32, BF = 48, BE = 64, BW = 80;  // BW: bulk copy of a slot's parameter vector",
    "  const uint32_t BX = 0, BA = 24, BD = 48, BF = 72, BE = 96, BW = 120;  // 8 bytes per tile slot each; BW: bulk copy of a slot's parameter vector")
sub("    for (int s = 0; s < 2; ++s) {\n      mbar_init(bars + BX + 8 * s, 1);", "    for (int s = 0; s < NSLOT; ++s) {\n      mbar_init(bars + BX + 8 * s, 1);")
sub("  uint32_t ph_x0 = 0, ph_x1 = 0, ph_a = 0, ph_d0 = 0, ph_d1 = 0, ph_f0 = 0, ph_f1 = 0, ph_e = 0, ph_w = 0;",
    "  uint32_t ph_x = 0, ph_d = 0, ph_f = 0;  // one parity bit per tile slot\n  uint32_t ph_a = 0, ph_e = 0, ph_w = 0;")
# ---------------------------------------------------------------- control warps
sub("      for (int t = s; t < n_tiles; t += 2) {", "      for (int t = s; t < n_tiles; t += NSLOT) {")
sub("          if (t >= 2 && l == (s == 0 ? 0 : L - 1)) {", "          if (t >= NSLOT && l == (s == 0 ? 0 : L - 1)) {")
sub("          const uint32_t dcol = (s == 1 && l == L - 1) ? tmem + COL_DX : tb + COL_D;",
    "          const uint32_t dcol = (s >= 1 && l == L - 1) ? tmem + COL_DX + (uint32_t)(s - 1) * 64u : tb + COL_D;")
sub("            if (l == 1 && t + 2 < n_tiles) {  // layer 0's MMAs (which read the x boxes) are complete => the boxes are free",
    "            if (l == 1 && t + NSLOT < n_tiles) {  // layer 0's MMAs (which read the x boxes) are complete => the boxes are free")
sub("              tma_load_2d(xdst, &map_x, 0, (int)(xrow0 + (long)(t + 2) * TILE), bar_x);\n              tma_load_2d(xdst + BOX_BYTES, &map_x, 32, (int)(xrow0 + (long)(t + 2) * TILE), bar_x);",
    "              tma_load_2d(xdst, &map_x, 0, (int)(xrow0 + (long)(t + NSLOT) * TILE), bar_x);\n              tma_load_2d(xdst + BOX_BYTES, &map_x, 32, (int)(xrow0 + (long)(t + NSLOT) * TILE), bar_x);")
# ---------------------------------------------------------------- layer warps
sub("      for (int t0 = 0; t0 < n_tiles; t0 += 2) {\n        // D -> bias, tanh -> next layer's A operand",
    "      for (int t0 = 0; t0 < n_tiles; t0 += NSLOT) {\n        // D -> bias, tanh -> next layer's A operand")
sub("          for (int s = 0; s < 2; ++s) {\n            if (t0 + s >= n_tiles) continue;", "          for (int s = 0; s < NSLOT; ++s) {\n            if (t0 + s >= n_tiles) continue;")
sub("            mbar_wait(bars + BD + 8 * s, s ? ph_d1 : ph_d0);\n            if (s) ph_d1 ^= 1; else ph_d0 ^= 1;",
    "            mbar_wait(bars + BD + 8 * s, (ph_d >> s) & 1u);\n            ph_d ^= 1u << s;")
# ---------------------------------------------------------------- output warps
sub("      const uint32_t ybox = stage + OBOX_BYTES;                                           // y rows of the second tile of a pair\n", "")
sub("      const uint32_t stage = sbase + a.stage_ofs + (warp - MAIN_WARPS) * 2 * OBOX_BYTES;  // transpose staging of the accumulator",
    "      const uint32_t stage = sbase + a.stage_ofs + (warp - MAIN_WARPS) * OBOX_BYTES;  // transpose staging of the accumulator")
sub("        const int s = tt & 1;\n        const uint32_t xbox = sbase + a.xbox_ofs + (2 * s + h) * BOX_BYTES + (uint32_t)row * 128u;\n        mbar_wait(bars + BX + 8 * s, s ? ph_x1 : ph_x0);\n        if (s) ph_x1 ^= 1; else ph_x0 ^= 1;",
    "        const int s = tt % NSLOT;\n        const uint32_t xbox = sbase + a.xbox_ofs + (2 * s + h) * BOX_BYTES + (uint32_t)row * 128u;\n        mbar_wait(bars + BX + 8 * s, (ph_x >> s) & 1u);\n        ph_x ^= 1u << s;")
sub("        const uint32_t sl = lane_base + (s == 1 ? COL_DX : COL_D) + h * 32;", "        const uint32_t sl = lane_base + (s >= 1 ? COL_DX + (uint32_t)(s - 1) * 64u : COL_D) + h * 32;")
# emit(): y always comes from registers
sub("      auto emit = [&](int t, bool y_smem) {", "      auto emit = [&](int t) {")
sub('          if (y_smem) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(yv.x), "=f"(yv.y), "=f"(yv.z), "=f"(yv.w) : "r"(ybox + (uint32_t)r * 128u + (uint32_t)tc * 16u));\n', "")
sub("trace_ev(a, ring, trace_cnt, 12, t, L - 1, t & 1);", "trace_ev(a, ring, trace_cnt, 12, t, L - 1, t % NSLOT);")
sub("trace_ev(a, ring, trace_cnt, 9, t, L - 1, t & 1);", "trace_ev(a, ring, trace_cnt, 9, t, L - 1, t % NSLOT);")
# the pair loop becomes a triple loop
a = src.index("      split_x(0);\n      if (n_tiles > 1) split_x(1);")
b = src.index("    fence_proxy_async();  // this item's generic accesses to the x boxes / staging precede the next item's bulk copy and TMA loads")
src = src[:a] + '''      // y rows of tile t -> registers (transposed layout), requested as early as the registers are free
      auto load_y = [&](int t) {
        const int trow = row_begin + t * TILE;
        const int nrows = min(TILE, row_end - trow);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = min(q * 32 + i * 4 + tr, nrows - 1);
          yt[i] = in_cols ? __ldg(reinterpret_cast<const float4*>(a.y + (job.x_row + trow + r) * (long)TP + h * 32) + tc) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      auto wait_f = [&](int s, int t) {
        mbar_wait(bars + BF + 8 * s, (ph_f >> s) & 1u);
        ph_f ^= 1u << s;
        tc_fence_after();
        if (lane == 0 && (warp == MAIN_WARPS || warp == EPI_WARPS - 1)) trace_ev(a, ring, trace_cnt, 8, t, L - 1, s);
      };

      split_x(0);
      if (n_tiles > 1) split_x(1);
      if (n_tiles > 2) split_x(2);

      for (int t0 = 0; t0 < n_tiles; t0 += NSLOT) {
        const int n_in = min(NSLOT, n_tiles - t0);
        if (has_y) load_y(t0);  // requested before the tile's accumulator is ready
        // ---- first what the layer pipeline waits for: free slot 0's accumulator, feed every slot its next tile
        if (lane == 0 && (warp == MAIN_WARPS || warp == EPI_WARPS - 1)) trace_ev(a, ring, trace_cnt, 7, t0, L - 1, 0);
        wait_f(0, t0);
        park(0, t0);
        if (t0 + NSLOT < n_tiles) split_x(t0 + NSLOT);  // the output-layer MMA of tile t0 is complete: nothing reads slot 0's A operands
        if (n_in > 1) {
          wait_f(1, t0 + 1);
          if (t0 + NSLOT + 1 < n_tiles) split_x(t0 + NSLOT + 1);  // slots 1 and 2 keep their output in a spare accumulator until parked below
        }
        if (n_in > 2) {
          wait_f(2, t0 + 2);
          if (t0 + NSLOT + 2 < n_tiles) split_x(t0 + NSLOT + 2);
        }
        // ---- then the stores
        emit(t0);
        if (n_in > 1) {
          if (has_y) load_y(t0 + 1);
          park(1, t0 + 1);
          emit(t0 + 1);
        }
        if (n_in > 2) {
          if (has_y) load_y(t0 + 2);
          park(2, t0 + 2);
          emit(t0 + 2);
        }
      }
    }
''' + src[b:]
# ---------------------------------------------------------------- host side: three slots of x boxes, staging without the y boxes
sub("  a.xbox_ofs = ofs; ofs += 4 * BOX_BYTES;            // two tile slots x two 32-column boxes",
    "  a.xbox_ofs = ofs; ofs += 2 * NSLOT * BOX_BYTES;    // NSLOT tile slots x two 32-column boxes")
sub("  a.stage_ofs = ofs; ofs += OUT_WARPS * 2 * OBOX_BYTES;  // per output warp: a 32-row x 32-column staging box + a y box of the same shape",
    "  a.stage_ofs = ofs; ofs += OUT_WARPS * OBOX_BYTES;  // per output warp: a 32-row x 32-column staging box")
sub("  GB_REQUIRE(a.param_bytes <= 4 * BOX_BYTES + OUT_WARPS * 2 * OBOX_BYTES, GB_E_SMEM,",
    "  GB_REQUIRE(a.param_bytes <= 2 * NSLOT * BOX_BYTES + OUT_WARPS * OBOX_BYTES, GB_E_SMEM,")
open(os.path.join(HERE, "ffae_infer_tc_v18_three_slots.cu"), "w").write(src)
print("written")
