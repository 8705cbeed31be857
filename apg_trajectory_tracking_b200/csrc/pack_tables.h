// Host side of the weight packing: the segment tables apg_pack_kernel (misc_kernels.cu) walks to turn the torch-flat
// parameter vector into the kernels' forward ([in][out], swizzled for the mma path) and backward ([out][in]) images.
// Header-only so that the CPU model of the kernels (tests/hostcheck) packs with the very same tables.
#pragma once
#include "layouts.h"

namespace apg {

// contiguous source [rows][cols]; destination row stride ldd, zero-filled up to ldd
inline void add_seg(PackTable& t, int which, int mode, int src, int dst, int rows, int cols, int ldd) {
  PackSeg& s = t.seg[t.n++];
  s.which = which; s.mode = mode; s.src = src; s.sld = cols; s.dst = dst; s.rows = rows; s.cols = cols;
  s.wcols = ldd; s.ldd = ldd; s.sw = 0; s.perm = 0;
}
// weight matrix read as mma B fragments: swizzled when its row stride is a multiple of 32 floats
inline void add_seg_mma(PackTable& t, int which, int mode, int src, int dst, int rows, int cols, int ldd) {
  add_seg(t, which, mode, src, dst, rows, cols, ldd);
  t.seg[t.n - 1].sw = mma_sw(ldd);
}
// general form: strided source, destination window of wcols columns inside rows of stride ldd
inline void add_seg_ex(PackTable& t, int which, int mode, int src, int sld, int dst, int rows, int cols, int wcols, int ldd) {
  PackSeg& s = t.seg[t.n++];
  s.which = which; s.mode = mode; s.src = src; s.sld = sld; s.dst = dst; s.rows = rows; s.cols = cols;
  s.wcols = wcols; s.ldd = ldd; s.sw = 0; s.perm = 0;
}

inline PackTable hutter_pack_table(const HutterLayout& y) {
  PackTable t;
  t.n = 0;
  // forward: [in][out]
  const int f0p8 = (y.F0 + 7) & ~7, krp8 = (y.KR + 7) & ~7;
  add_seg_mma(t, 0, PK_TRANSPOSE, y.t_ws, y.f_ws, HID, y.F0, HID);
  if (f0p8 > y.F0) add_seg_ex(t, 0, PK_COPY_PAD, y.t_ws, 1, y.f_ws + y.F0 * HID, f0p8 - y.F0, 0, HID, HID);   // zero rows
  add_seg(t, 0, PK_COPY_PAD, y.t_bs, y.f_bs, 1, HID, HID);
  if (y.conv) {
    add_seg(t, 0, PK_CONV_FWD, y.t_wc, y.f_wr, CONV_CH, y.KC, 24);
    if (krp8 > y.KR) add_seg_ex(t, 0, PK_COPY_PAD, y.t_wc, 1, y.f_wr + y.KR * 24, krp8 - y.KR, 0, 24, 24);
    add_seg(t, 0, PK_COPY_PAD, y.t_bc, y.f_br, 1, CONV_CH, CONV_CH);
  } else {
    add_seg_mma(t, 0, PK_TRANSPOSE, y.t_wr, y.f_wr, HID, y.LR, HID);
    if (krp8 > y.KR) add_seg_ex(t, 0, PK_COPY_PAD, y.t_wr, 1, y.f_wr + y.KR * HID, krp8 - y.KR, 0, HID, HID);
    add_seg(t, 0, PK_COPY_PAD, y.t_br, y.f_br, 1, HID, HID);
  }
  add_seg_mma(t, 0, PK_TRANSPOSE, y.t_w1, y.f_w1, HID, y.K1, HID);
  t.seg[t.n - 1].perm = y.perm_npos;
  add_seg(t, 0, PK_COPY_PAD, y.t_b1, y.f_b1, 1, HID, HID);
  add_seg_mma(t, 0, PK_TRANSPOSE, y.t_w2, y.f_w2, HID, HID, HID);
  add_seg(t, 0, PK_COPY_PAD, y.t_b2, y.f_b2, 1, HID, HID);
  add_seg_mma(t, 0, PK_TRANSPOSE, y.t_w3, y.f_w3, HID, HID, HID);
  add_seg(t, 0, PK_COPY_PAD, y.t_b3, y.f_b3, 1, HID, HID);
  add_seg_mma(t, 0, PK_TRANSPOSE, y.t_wo, y.f_wo, y.Mo, HID, y.ld_fwo);
  add_seg(t, 0, PK_COPY_PAD, y.t_bo, y.f_bo, 1, y.Mo, y.Mo4);
  // backward: [out][in]
  add_seg_mma(t, 1, PK_COPY_PAD, y.t_wo, y.b_wo, y.Mo, HID, HID);
  add_seg_mma(t, 1, PK_COPY_PAD, y.t_w3, y.b_w3, HID, HID, HID);
  add_seg_mma(t, 1, PK_COPY_PAD, y.t_w2, y.b_w2, HID, HID, HID);
  add_seg_mma(t, 1, PK_COPY_PAD, y.t_w1, y.b_w1, HID, y.K1, y.ld_bw1);
  t.seg[t.n - 1].perm = y.perm_npos;
  add_seg(t, 1, PK_COPY_PAD, y.t_ws, y.b_ws, HID, y.F0, y.ld_bws);
  if (y.conv) add_seg(t, 1, PK_CONV_BWD, y.t_wc, y.b_wr, CONV_CH, y.KC, y.ld_bwr);
  else        add_seg(t, 1, PK_COPY_PAD, y.t_wr, y.b_wr, HID, y.LR, y.ld_bwr);
  return t;
}

inline PackTable lstm_pack_table(const LstmLayout& y) {
  PackTable t;
  t.n = 0;
  const int G = 4 * LSTM_HS, f0p = pad4(y.F0), nc = CONV_CH * y.npos, mo4 = pad4(y.Mo);
  add_seg(t, 0, PK_CONV_FWD, y.t_wc, y.f_wc, CONV_CH, y.KC, 24);
  if (((y.KC + 7) & ~7) > y.KC)
    add_seg_ex(t, 0, PK_COPY_PAD, y.t_wc, 1, y.f_wc + y.KC * 24, ((y.KC + 7) & ~7) - y.KC, 0, 24, 24);
  add_seg(t, 0, PK_COPY_PAD, y.t_bc, y.f_bc, 1, CONV_CH, CONV_CH);
  add_seg_ex(t, 0, PK_TRANSPOSE, y.t_wih, y.IH, y.f_wg, G, y.F0, G, G);
  if (f0p > y.F0) add_seg_ex(t, 0, PK_COPY_PAD, y.t_wih, 1, y.f_wg + y.F0 * G, f0p - y.F0, 0, G, G);
  add_seg_ex(t, 0, PK_TRANSPOSE, y.t_wih + y.F0, y.IH, y.f_wg + f0p * G, G, nc, G, G);
  add_seg_ex(t, 0, PK_TRANSPOSE, y.t_whh, LSTM_HS, y.f_wg + y.KX * G, G, LSTM_HS, G, G);
  add_seg(t, 0, PK_COPY_PAD, y.t_bih, y.f_bih, 1, G, G);
  add_seg(t, 0, PK_COPY_PAD, y.t_bhh, y.f_bhh, 1, G, G);
  add_seg(t, 0, PK_TRANSPOSE, y.t_wo, y.f_wo, y.Mo, LSTM_HS, mo4);
  add_seg(t, 0, PK_COPY_PAD, y.t_bo, y.f_bo, 1, y.Mo, mo4);
  add_seg_ex(t, 1, PK_COPY_PAD, y.t_wih, y.IH, y.b_wg, G, y.F0, f0p, y.KG);
  add_seg_ex(t, 1, PK_COPY_PAD, y.t_wih + y.F0, y.IH, y.b_wg + f0p, G, nc, nc, y.KG);
  add_seg_ex(t, 1, PK_COPY_PAD, y.t_whh, LSTM_HS, y.b_wg + y.KX, G, LSTM_HS, LSTM_HS, y.KG);
  add_seg(t, 1, PK_COPY_PAD, y.t_wo, y.b_wo, y.Mo, LSTM_HS, LSTM_HS);
  add_seg(t, 1, PK_CONV_BWD, y.t_wc, y.b_wc, CONV_CH, y.KC, y.ld_bwr);
  return t;
}

inline PackTable simple_pack_table(const SimpleLayout& y) {
  PackTable t;
  t.n = 0;
  for (int l = 0; l < SIMPLE_NL; ++l) {
    add_seg(t, 0, PK_TRANSPOSE, y.t_w[l], y.f_w[l], y.dout[l], y.din[l], y.ldf[l]);
    add_seg(t, 0, PK_COPY_PAD, y.t_b[l], y.f_b[l], 1, y.dout[l], y.ldf[l]);
    add_seg(t, 1, PK_COPY_PAD, y.t_w[l], y.b_w[l], y.dout[l], y.din[l], y.ldb[l]);
  }
  return t;
}

}  // namespace apg
