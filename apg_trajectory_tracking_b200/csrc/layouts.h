// Parameter / shared-memory layouts of the policy networks, shared by host (capi.cu) and device code.
//
// "torch flat" = the tensors of net.parameters() concatenated in order, each in its torch layout.  This is what the
// C-ABI takes (params) and returns (grad_params).
//   hutter (models/hutter_model.py:12-30): states_in.{w[64][F0],b}, conv_ref.{w[20][RD][3],b}, ref_in.{w[64][L*RD],b},
//                                          fc1.{w[64][K1],b}, fc2, fc3, fc_out.{w[Mo][64],b}
//   simple (models/simple_model.py:9-18) : fc0.{w[32][4],b}, fc1.{w[64][32],b}, fc2.{w[64][64],b}, fc3.{w[32][64],b},
//                                          fc_out.{w[Mo][32],b}
//   lstm   (models/rnn.py:7-27)          : conv_ref.{w,b}, ref_in.{w,b}, fc_out.{w[Mo][8],b},
//                                          lstm.{weight_ih[32][F0+20*npos], weight_hh[32][8], bias_ih[32], bias_hh[32]}
// "packed fwd" = [in][out] (transposed, output dim padded to a multiple of 4) for the forward GEMMs;
// "packed bwd" = [out][in] (input dim padded to a multiple of 4) for the dX GEMMs.  Both are produced on the device
// by apg_pack_kernel once per call from the torch-flat vector.
#pragma once

// Kernel launch.  CUDA: expands to exactly `kernel<<<grid, block, smem, stream>>>` (the kernel name comes last so that
// template arguments with commas survive the preprocessor).  -DAPG_SIM (tests only): every CTA of the grid runs on
// the CPU model of tests/hostcheck/gpu_sim.h, one OS thread per GPU thread.
#ifdef APG_SIM
#define APG_LAUNCH(grid, block, smem, stream, ...) \
  ::sim::Launcher((grid), (block), (size_t)(smem)).bind([&](auto&&... sim_args_) { __VA_ARGS__(sim_args_...); })
#define APG_LAUNCH_PDL APG_LAUNCH
#else
#define APG_LAUNCH(grid, block, smem, stream, ...) __VA_ARGS__<<<(grid), (block), (smem), (stream)>>>
#ifdef __CUDACC__
// The same launch with programmatic stream serialization: the kernel's CTAs may start while the preceding kernel of
// the stream is still draining (as its CTAs leave their SMs); the kernel itself executes griddepcontrol.wait
// (tcp::griddep_wait) before it touches anything the preceding kernels wrote.  What it buys here: the 5-7 us between
// the end of one 148-CTA kernel and the first instruction of the next (measured with %globaltimer, profiles/r2).
namespace apg {
template <typename... KArgs>
struct PdlLaunch {
  void (*kernel)(KArgs...);
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  PdlLaunch(void (*k)(KArgs...), int grid, int block, size_t smem, cudaStream_t st) : kernel(k), cfg{} {
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  template <typename... Args>
  void operator()(Args&&... args) { (void)cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...); }
};
template <typename... KArgs>
PdlLaunch<KArgs...> pdl_launch(void (*k)(KArgs...), int grid, int block, size_t smem, cudaStream_t st) {
  return PdlLaunch<KArgs...>(k, grid, block, smem, st);
}
}  // namespace apg
#define APG_LAUNCH_PDL(grid, block, smem, stream, ...) ::apg::pdl_launch(__VA_ARGS__, (grid), (block), (smem), (stream))
#endif
#endif

namespace apg {

enum NetKind { NET_HUTTER_CONV = 0, NET_HUTTER_LIN = 1, NET_SIMPLE = 2, NET_LSTM = 3 };
enum Mode { MODE_CONCURRENT = 0, MODE_AUTOREGRESSIVE = 1, MODE_LSTM = 2 };
enum Window { WINDOW_CUMULATIVE = 0, WINDOW_RELATIVE = 1 };

constexpr int TM = 64;        // drones per tile
constexpr int TMP = 68;       // padded row length of feature-major tiles (floats)
constexpr int NT = 256;       // threads per CTA
constexpr int NWARP = NT / 32;

inline __host__ __device__ int pad4(int x) { return (x + 3) & ~3; }

// Leading dimension of a weight matrix whose rows are read as mma.sync B fragments (lane (g,t) reads row k0+t,
// column n0+g): rows must land in distinct 8-bank groups.  ld % 32 == 8 or 24 does that by itself; ld % 32 == 0
// needs the XOR swizzle col ^ ((row & 3) << 3)  (mma_sw); ld % 32 == 16 is padded by 8.
inline __host__ __device__ int mma_ld(int n) {
  int l = (n + 7) & ~7;
  if ((l & 31) == 16) l += 8;
  return l;
}
inline __host__ __device__ int mma_sw(int ld) { return (ld & 31) == 0; }

constexpr int HID = 64;       // hidden width of the hutter nets
constexpr int CONV_CH = 20;   // conv_ref output channels

struct HutterLayout {
  int F0, L, RD, Mo, conv;      // state features, reference rows seen by the net, reference width, outputs, conv?
  int npos, KC, LR, KR, NRtot, K1, Mo4, XR, CR;   // CR: rows of the d-logit buffer of the adjoint kernel
  // torch flat offsets
  int t_ws, t_bs, t_wc, t_bc, t_wr, t_br, t_w1, t_b1, t_w2, t_b2, t_w3, t_b3, t_wo, t_bo, n_params;
  // packed forward
  int f_ws, f_bs, f_wr, f_br, f_w1, f_b1, f_w2, f_b2, f_w3, f_b3, f_wo, f_bo, f_total, ld_wr;
  int ld_fwo;                   // row stride of the packed fc_out forward weights (mma_ld(Mo))
  // packed backward ([out][in_padded])
  int b_wo, b_w3, b_w2, b_w1, b_ws, b_wr, b_total, ld_bws, ld_bwr;
  int ld_bw1;                   // row stride of the packed fc1 backward weights (mma_ld(K1))
  // row of conv output (channel c, position t) inside the activation arena, relative to the first conv row:
  // c*conv_cs + t*conv_ts.  The reference flattens channel-major (cs = npos, ts = 1); the hutter kernels keep the
  // tile position-major (cs = 1, ts = 20: consecutive channels in consecutive rows -> conflict-free mma fragments)
  // and permute the fc1 weight columns at pack time (perm_npos).
  int conv_cs, conv_ts, perm_npos;
};

inline __host__ HutterLayout make_hutter_layout(int F0, int L, int RD, int Mo, int conv) {
  HutterLayout y;
  y.F0 = F0; y.L = L; y.RD = RD; y.Mo = Mo; y.conv = conv;
  y.npos = conv ? L - 2 : 1;
  y.KC = 3 * RD;
  y.LR = L * RD;
  y.KR = conv ? y.KC : y.LR;
  y.NRtot = conv ? CONV_CH * y.npos : HID;
  y.K1 = HID + y.NRtot;
  y.Mo4 = pad4(Mo);
  y.conv_cs = 1; y.conv_ts = CONV_CH; y.perm_npos = conv ? y.npos : 0;
  y.XR = y.K1 > HID + y.Mo4 ? y.K1 : HID + y.Mo4;
  y.CR = y.Mo4 > HID ? y.Mo4 : HID;
  int o = 0;
  y.t_ws = o; o += HID * F0;      y.t_bs = o; o += HID;
  y.t_wc = o; o += CONV_CH * RD * 3; y.t_bc = o; o += CONV_CH;
  y.t_wr = o; o += HID * y.LR;    y.t_br = o; o += HID;
  y.t_w1 = o; o += HID * y.K1;    y.t_b1 = o; o += HID;
  y.t_w2 = o; o += HID * HID;     y.t_b2 = o; o += HID;
  y.t_w3 = o; o += HID * HID;     y.t_b3 = o; o += HID;
  y.t_wo = o; o += Mo * HID;      y.t_bo = o; o += Mo;
  y.n_params = o;
  const int nr = conv ? CONV_CH : HID;
  o = 0;
  // first-layer weights feed mma B fragments: K rows padded to 8 (zero rows), conv columns padded to 24
  const int ldr = conv ? 24 : HID;
  y.ld_wr = ldr;
  y.f_ws = o; o += ((F0 + 7) & ~7) * HID;       y.f_bs = o; o += HID;
  y.f_wr = o; o += ((y.KR + 7) & ~7) * ldr;     y.f_br = o; o += pad4(nr);
  y.f_w1 = o; o += y.K1 * HID;          y.f_b1 = o; o += HID;
  y.f_w2 = o; o += HID * HID;           y.f_b2 = o; o += HID;
  y.f_w3 = o; o += HID * HID;           y.f_b3 = o; o += HID;
  y.ld_fwo = mma_ld(y.Mo4);
  y.f_wo = o; o += HID * y.ld_fwo;      y.f_bo = o; o += y.Mo4;
  y.f_total = o;
  y.ld_bws = pad4(F0);
  y.ld_bwr = pad4(y.KR);
  o = 0;
  y.b_wo = o; o += Mo * HID;
  y.b_w3 = o; o += HID * HID;
  y.b_w2 = o; o += HID * HID;
  y.ld_bw1 = mma_ld(y.K1);
  y.b_w1 = o; o += HID * y.ld_bw1;
  y.b_ws = o; o += HID * y.ld_bws;
  y.b_wr = o; o += nr * y.ld_bwr;
  y.b_total = o;
  return y;
}

// ---- simple (cartpole) MLP: 4 -> 32 -> 64 -> 64 -> 32 -> Mo, tanh everywhere (models/simple_model.py:9-28)
constexpr int SIMPLE_NL = 5;
struct SimpleLayout {
  int F0, Mo, Mo4;
  int din[SIMPLE_NL], dout[SIMPLE_NL];          // real layer dims
  int t_w[SIMPLE_NL], t_b[SIMPLE_NL], n_params;  // torch flat offsets
  int f_w[SIMPLE_NL], f_b[SIMPLE_NL], ldf[SIMPLE_NL], f_total;   // packed fwd [in][out4]
  int b_w[SIMPLE_NL], ldb[SIMPLE_NL], b_total;                   // packed bwd [out][in4]
  int row[SIMPLE_NL], rows_total;               // activation arena: row offset of each layer's output
};

inline __host__ SimpleLayout make_simple_layout(int F0, int Mo) {
  SimpleLayout y;
  y.F0 = F0; y.Mo = Mo; y.Mo4 = pad4(Mo);
  const int dims[SIMPLE_NL + 1] = {F0, 32, 64, 64, 32, Mo};
  int ot = 0, of = 0, ob = 0, r = 0;
  for (int l = 0; l < SIMPLE_NL; ++l) {
    y.din[l] = dims[l]; y.dout[l] = dims[l + 1];
    y.t_w[l] = ot; ot += dims[l + 1] * dims[l];
    y.t_b[l] = ot; ot += dims[l + 1];
    y.ldf[l] = pad4(dims[l + 1]);
    y.f_w[l] = of; of += dims[l] * y.ldf[l];
    y.f_b[l] = of; of += y.ldf[l];
    y.ldb[l] = pad4(dims[l]);
    y.b_w[l] = ob; ob += dims[l + 1] * y.ldb[l];
    y.row[l] = r; r += y.ldf[l];
  }
  y.n_params = ot; y.f_total = of; y.b_total = ob; y.rows_total = r;
  return y;
}

// ---- LSTM policy (models/rnn.py): conv encoder -> LSTMCell(15 + 20*npos, 8) -> Linear(8, 4)
// Per-step activation arena of a tile (feature-major rows of TMP floats), stashed whole for the adjoint:
//   [0,15) features, 15 zero pad, [16,KX) conv (c*npos+t), [KX,KX+8) h_prev, then c_prev(8), gates i|f|g|o (32),
//   c'(8), h'(8), actions(4)
constexpr int LSTM_HS = 8;
struct LstmLayout {
  int F0, L, RD, Mo, npos, KC, LR, IH, KX, KG, ROWS;
  int R_HP, R_CP, R_G, R_C, R_H, R_A;
  int t_wc, t_bc, t_wr, t_br, t_wo, t_bo, t_wih, t_whh, t_bih, t_bhh, n_params;
  int f_wc, f_bc, f_wg, f_bih, f_bhh, f_wo, f_bo, f_total;
  int b_wg, b_wo, b_wc, ld_bwr, b_total;
  HutterLayout cv;     // the conv-encoder fields the shared conv helpers read (npos, KC, LR, RD, L, t_wc, t_bc, ld_bwr)
};

inline __host__ LstmLayout make_lstm_layout(int F0, int L, int RD, int Mo) {
  LstmLayout y;
  y.F0 = F0; y.L = L; y.RD = RD; y.Mo = Mo;
  y.npos = L - 2; y.KC = 3 * RD; y.LR = L * RD;
  y.IH = F0 + CONV_CH * y.npos;
  y.KX = pad4(F0) + CONV_CH * y.npos;
  y.KG = y.KX + LSTM_HS;
  y.R_HP = y.KX; y.R_CP = y.KX + 8; y.R_G = y.KX + 16; y.R_C = y.KX + 48; y.R_H = y.KX + 56; y.R_A = y.KX + 64;
  y.ROWS = y.KX + 68;
  int o = 0;
  y.t_wc = o; o += CONV_CH * RD * 3;  y.t_bc = o; o += CONV_CH;
  y.t_wr = o; o += HID * y.LR;        y.t_br = o; o += HID;
  y.t_wo = o; o += Mo * LSTM_HS;      y.t_bo = o; o += Mo;
  y.t_wih = o; o += 4 * LSTM_HS * y.IH;
  y.t_whh = o; o += 4 * LSTM_HS * LSTM_HS;
  y.t_bih = o; o += 4 * LSTM_HS;
  y.t_bhh = o; o += 4 * LSTM_HS;
  y.n_params = o;
  o = 0;
  y.f_wc = o; o += ((y.KC + 7) & ~7) * 24;  y.f_bc = o; o += CONV_CH;
  y.f_wg = o; o += y.KG * 4 * LSTM_HS;
  y.f_bih = o; o += 4 * LSTM_HS;          y.f_bhh = o; o += 4 * LSTM_HS;
  y.f_wo = o; o += LSTM_HS * pad4(Mo);    y.f_bo = o; o += pad4(Mo);
  y.f_total = o;
  y.ld_bwr = pad4(y.KC);
  o = 0;
  y.b_wg = o; o += 4 * LSTM_HS * y.KG;
  y.b_wo = o; o += pad4(Mo * LSTM_HS);
  y.b_wc = o; o += CONV_CH * y.ld_bwr;
  y.b_total = o;
  y.cv = make_hutter_layout(F0, L, RD, Mo, 1);
  y.cv.t_wc = y.t_wc; y.cv.t_bc = y.t_bc; y.cv.ld_bwr = y.ld_bwr;
  y.cv.conv_cs = y.npos; y.cv.conv_ts = 1; y.cv.perm_npos = 0;      // the LSTM arena stays channel-major
  return y;
}

// Segment table of the pack kernel: dst[...] = src[...] with a layout transform.
enum PackMode { PK_COPY_PAD = 0, PK_TRANSPOSE = 1, PK_CONV_FWD = 2, PK_CONV_BWD = 3 };
// src is [rows][cols] with row stride sld; the destination window is `wcols` wide (zero-filled beyond the data)
// with row stride ldd.  which: 0 -> fwd buffer, 1 -> bwd buffer
// sw: XOR-swizzle the dst columns; perm: fc1 input index (position-major, see HutterLayout) -> torch column
struct PackSeg { int src, sld, dst, rows, cols, wcols, ldd, mode, which, sw, perm; };
constexpr int MAX_PACK_SEGS = 24;
struct PackTable { int n; PackSeg seg[MAX_PACK_SEGS]; };

}  // namespace apg
