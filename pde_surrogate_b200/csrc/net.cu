// net.cu — DenseED executor: builds the layer list of the reference's DenseED
// (models/codec.py:211-293), owns the workspace layout, and runs forward / backward as a fixed
// sequence of kernel launches on the caller's stream.  Also the single-convolution C-ABI entry
// points used by the unit tests.
//
// Design (B200-first, not a translation of nn.Sequential):
//  * activations are NHWC; every dense block lives in ONE preallocated buffer and each layer
//    writes its growth_rate-channel slice (no torch.cat copies, codec.py:73-75);
//  * BatchNorm+ReLU never materialise: the producing conv's epilogue accumulates per-channel
//    sum / sum-of-squares (fp64 atomics) once, every consumer folds its own gamma/beta into the
//    load of its operand;
//  * nearest x2 upsampling (codec.py:24-30) is an addressing mode of the consumer conv;
//  * backward: the dgrad epilogue applies ReLU mask + BatchNorm backward and accumulates
//    scale*dZ into the block's gradient buffer; the per-channel mean corrections of all
//    consumers are applied lazily, once, to the 16-channel slice a producer reads as its dY
//    (fix_dy), so no dZ tensor is ever written.
#include <string>
#include <vector>
#include <string.h>
#include <stdlib.h>
#include "conv.cuh"
#include "conv_tc.cuh"
#include "first_conv.cuh"

namespace pdes {

struct Buf {
  int H, W, C, ld;
  size_t act, grad;  // float offsets into the workspace
  size_t stat;       // double offset: sum[C], sumsq[C]
};

struct Layer {
  int kind;  // 0 plain conv (In_conv), 1 dense layer, 2 BN-ReLU-conv of a transition / last decoding
  std::string conv_name, bn_name;
  int in_buf, out_buf, coff;
  int Cin, Cout, KS, stride, pad, up;
  int Hs, Ws, Ho, Wo;
  int64_t w_off, g_off, b_off, rm_off, rv_off;
  size_t wf, wb;  // float offsets of the packed weights
  size_t bsum;    // double offset
  int CinP, CoP, CoutPb, CiPb;
  bool last_consumer;
  // tcgen05 path
  int Nf, Nb;        // padded GEMM-N (Cout / Cin rounded up to 16)
  bool tc2_fwd, tc2_bwd;  // TMA-fed two-piece fp16 tcgen05 path (conv_tc2.cu)
  Tc2Plan p2f, p2b;
  size_t w2f, w2b;   // float offsets of the packed fp16 filter pieces
  size_t planes;     // float offset of this layer's activation piece planes [2][B][Hv][Wv][Cp] fp16
  size_t planesB;    // float offset of this layer's dY piece planes (read by its dgrad and, concurrently, its wgrad)
  bool tc_wg;        // weight gradient on tcgen05
  bool first_k;      // dedicated CUDA-core kernels of the first convolution (first_conv.cu)
  bool dense_fwd;    // fused BatchNorm+ReLU+split+conv forward of a thin 3x3 layer (conv_dense.cu)
  size_t wdn;        // float offset of its packed filter ("dx in N" layout)
  bool bilinear;     // the x2 upsampling in front of this convolution is materialised by bilinear.cu (align_corners
                     // interpolation, or zero insertion for the transposed convolution) instead of nearest
  bool no_drop;      // no nn.Dropout2d behind this convolution (conv1 of a bottleneck dense layer)
  bool dg_im2col;    // data gradient = ONE 1x1 GEMM over the (tap, co)-expanded dY planes the weight gradient builds
  Tc2Plan p2i;       // its conv_tc2 plan (KS = 1, GEMM-K = padded taps * Cout)
  size_t w2i = 0;    // float offset of its packed filter [(tap, co)][ci]
  bool convT;        // nn.ConvTranspose2d(k3, s2, p1, op1): zero-insert x2 + 3x3 conv with the flipped, transposed filter
  size_t wt = 0, gt = 0;   // float offsets (convT): equivalent Conv2d filter / its gradient
  bool drop;         // an nn.Dropout2d follows this convolution when the network has drop_rate > 0
  int drop_cprefix;  // channels of the dropout sites before this one (mask block offset = B * drop_cprefix)
  bool dense_bwd;    // fused dY-correction+split+dgrad of a thin 3x3 layer (conv_dense_bwd.cu)
  size_t wdb;        // float offset of its packed filter ("dx in K" layout)
  bool wg_taps_n;    // weight gradient as ONE 1x1 GEMM over an expanded dY (few output channels, DyIm2colArgs)
  size_t planesI;    // float offset of the expanded dY planes
  int ci_pad, co_pad;
  size_t dwp;        // float offset of the [tap][ci_pad][co_pad] staging gradient
};

struct ParamInfo {
  std::string name;
  int64_t offset;
  int ndim;
  int64_t shape[4];
  int kind;
};

}  // namespace pdes

using namespace pdes;

struct pdes_net {
  pdes_densenet_config cfg;
  std::vector<Buf> bufs;
  std::vector<Layer> layers;
  std::vector<ParamInfo> params;
  int64_t param_floats = 0, running_floats = 0;
  size_t ws_floats = 0, ws_doubles = 0, ws_bytes = 0, off_doubles = 0, off_tables = 0;
  size_t xin = 0;
  bool coupling = false;    // arch 2: cGlow _DenseCoupling (no In_conv: the input IS the block's first channels)
  int64_t zb_off = -1, zs_off = -1;   // Conv2dZeros bias / scale in the flat parameter buffer
  size_t out_keep = 0, dyg = 0;       // float offsets: forward output copy / dout * gain (arch 2)
  size_t up_scratch = 0;   // float offset: fp32 NHWC scratch of the bilinear-upsampling layers (a_up / dA_up)
  int in_hw = 0, out_hw = 0;  // spatial size of the network input / output (DenseED: both imsize)
  size_t gmax_off = 0;  // double offset: running |G| maxima (unsigned float bits), one per buffer + one for dout
  size_t dyinv = 0;     // float offset: per-layer inverse of the dynamic dY scale (written by the dY split)
  // side streams for the weight-gradient kernels: they are off the critical path of the backward
  // pass (nothing but the final unpack reads them), so they overlap with the dgrad chain
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
  int n_side = 0, n_side_req = 1, side_dev = -1;
  int n_bn = 0, maxC = 0, max_pack = 0;
  int n_wg = 0, n_tc2 = 0;
  size_t max_tc2_pack = 0;
  int max_wg_elems = 0, max_wg_cin = 0;  // largest unpack slab (Cout * 8 * taps) / Cin over the wgrad layers
  // executor-level CUDA graphs: the launch sequence of one forward / backward at a given batch size is
  // captured once (second call) and replayed afterwards; static input/output staging keeps pointers stable
  struct GraphSlot {
    int B = 0, training = -1, calls = 0, launches = 0;
    cudaGraphExec_t exec = nullptr;
  };
  GraphSlot gfwd[4], gbwd[2];
  int use_graph = 0;
  cudaStream_t cap_stream = nullptr;   // private stream the executor graphs are captured on
  size_t xs = 0, outs = 0, douts = 0;  // float offsets of the static x / out / dout buffers
  int n_wg_bound = 0;
  int tc_mask = 7;  // bit 0: forward, bit 1: dgrad, bit 2: wgrad on tcgen05
  int dense_on = 1; // PDES_DENSE_FWD=0: thin layers go through operand split + conv_tc2 (round-1 path)
  int dense_bwd_on = 1;  // PDES_DENSE_BWD=0: thin-layer dgrad through dY split + conv_tc2
  const float* drop_masks = nullptr;  // masks of the current training pass (pdes_densenet_set_dropout)
  int lowp = 0;     // PDES_CONV_DTYPE = fp32 (default: two fp16 pieces, three products) | fp16 | bf16 (one piece)
  // bound
  float* p = nullptr;
  float* g = nullptr;
  float* run = nullptr;
  unsigned char* ws = nullptr;
  int last_B = 0;
  bool fwd_train_done = false;
  int launches = 0;
  int conv_impl = 0;
  // PDES_TIMING=1: one CUDA event after every launch of the eager executor (diagnostics)
  bool timing = false;
  struct Mark {
    std::string first;
    cudaEvent_t second;
    double flops;
  };
  std::vector<Mark> marks;
};

namespace {

int64_t pad4(int64_t v) { return (v + 3) & ~(int64_t)3; }

bool stream_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return cs != cudaStreamCaptureStatusNone;
}

void mark(pdes_net* n, cudaStream_t st, const std::string& label, double flops = 0.0) {
  if (!n->timing) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  pdes_net::Mark m;
  m.first = label;
  m.second = e;
  m.flops = flops;
  n->marks.push_back(m);
}
int rup(int v, int m) { return (v + m - 1) / m * m; }

int add_buf(pdes_net* n, int H, int W, int C) {
  Buf b;
  b.H = H;
  b.W = W;
  b.C = C;
  b.ld = rup(C, 4);
  b.act = b.grad = b.stat = 0;
  n->bufs.push_back(b);
  return (int)n->bufs.size() - 1;
}

int conv_out(int in, int k, int s, int p) { return (in + 2 * p - k) / s + 1; }

void add_layer(pdes_net* n, int kind, const std::string& conv_name, const std::string& bn_name,
               int in_buf, int out_buf, int coff, int Cin, int Cout, int KS, int stride, int pad,
               int up, int Hs, int Ws) {
  Layer L;
  L.kind = kind;
  L.conv_name = conv_name;
  L.bn_name = bn_name;
  L.in_buf = in_buf;
  L.out_buf = out_buf;
  L.coff = coff;
  L.Cin = Cin;
  L.Cout = Cout;
  L.KS = KS;
  L.stride = stride;
  L.pad = pad;
  L.up = up;
  L.Hs = Hs;
  L.Ws = Ws;
  const int Hv = up ? 2 * Hs : Hs, Wv = up ? 2 * Ws : Ws;
  L.Ho = conv_out(Hv, KS, stride, pad);
  L.Wo = conv_out(Wv, KS, stride, pad);
  L.w_off = L.g_off = L.b_off = L.rm_off = L.rv_off = -1;
  L.wf = L.wb = L.bsum = 0;
  L.CinP = rup(Cin, 4);
  L.CoP = rup(Cout, 16);
  L.CoutPb = rup(Cout, 4);
  L.CiPb = rup(Cin, 16);
  L.last_consumer = false;
  L.Nf = rup(Cout, 16);
  L.Nb = rup(Cin, 16);
  L.tc2_fwd = L.tc2_bwd = false;
  L.w2f = L.w2b = L.planes = L.planesB = 0;
  memset(&L.p2f, 0, sizeof(L.p2f));
  memset(&L.p2b, 0, sizeof(L.p2b));
  L.tc_wg = false;
  L.first_k = false;
  L.wg_taps_n = false;
  L.dense_fwd = false;
  L.wdn = 0;
  L.dense_bwd = false;
  L.wdb = 0;
  L.drop = false;
  L.drop_cprefix = 0;
  L.bilinear = false;
  L.convT = false;
  L.dg_im2col = false;
  L.no_drop = false;
  memset(&L.p2i, 0, sizeof(L.p2i));
  L.planesI = 0;
  L.ci_pad = L.co_pad = 0;
  L.dwp = 0;
  n->layers.push_back(L);
}

// Mirrors DenseED.__init__ (models/codec.py:230-290) for the configuration the training script
// uses: dense layers without bottleneck, bottleneck transitions, nearest upsampling.
int build(pdes_net* n) {
  const pdes_densenet_config& c = n->cfg;
  PDES_REQUIRE(c.arch >= 0 && c.arch <= 2, PDES_ERR_INVALID,
               "pdes_densenet_create: arch %d (0 DenseED, 1 Decoder, 2 coupling network)", c.arch);
  PDES_REQUIRE(c.n_blocks >= 1 && c.n_blocks <= 15 && (c.arch >= 1 || c.n_blocks == 1 || c.n_blocks % 2 == 1),
               PDES_ERR_INVALID, "length of blocks must be an odd number, but got %d", c.n_blocks);
  PDES_REQUIRE(c.in_channels >= 1 && c.out_channels >= 1 && c.imsize >= (c.arch >= 1 ? 2 : 8) && c.growth_rate >= 1 &&
                   c.init_features >= 1 && c.max_batch >= 1,
               PDES_ERR_INVALID, "pdes_densenet_create: invalid configuration");
  for (int i = 0; i < c.n_blocks; ++i)
    PDES_REQUIRE(c.blocks[i] >= 1 && c.blocks[i] < kMaxConsumers - 1, PDES_ERR_UNSUPPORTED,
                 "dense block of %d layers (supported: 1..%d)", c.blocks[i], kMaxConsumers - 2);
  if (c.arch == 2) {
    // _DenseCoupling (models/glow_msc.py:276-294): `blocks[0]` dense layers on the input, then
    // reduce = BatchNorm -> ReLU -> Conv2dZeros(num_features, out_channels)
    PDES_REQUIRE(c.n_blocks == 1 && c.in_channels <= 256, PDES_ERR_INVALID,
                 "coupling network: one dense block, at most 256 input channels");
    n->coupling = true;
    n->in_hw = n->out_hw = c.imsize;
    const int H = c.imsize, C0 = c.in_channels, nl = c.blocks[0];
    const int cur = add_buf(n, H, H, C0 + nl * c.growth_rate);
    char nm[128];
    for (int j = 0; j < nl; ++j) {
      snprintf(nm, sizeof(nm), "denselayer%d", j + 1);
      add_layer(n, 1, std::string(nm) + ".conv1", std::string(nm) + ".norm1", cur, cur, C0 + j * c.growth_rate,
                C0 + j * c.growth_rate, c.growth_rate, 3, 1, 1, 0, H, H);
    }
    add_layer(n, 2, "reduce.conv_zero.conv", "reduce.norm1", cur, -1, 0, C0 + nl * c.growth_rate, c.out_channels, 3, 1,
              1, 0, H, H);
  } else {
  const bool decoder = c.arch == 1;
    const int n_enc = decoder ? 0 : c.n_blocks / 2;
    const int pad0 = (c.imsize % 2 == 0) ? 3 : 2;  // codec.py:238
    int H = decoder ? c.imsize : conv_out(c.imsize, 7, 2, pad0);
    int C = c.init_features;
    n->in_hw = c.imsize;
    // every dense block (or single-consumer tensor) is one buffer
    auto block_channels = [&](int C0, int nl) { return C0 + nl * c.growth_rate; };
    int cur = add_buf(n, H, H, block_channels(C, c.blocks[0]));
    if (decoder)  // Decoder.conv0: Conv2d(dim_latent, init_features, 3, 1, 1, bias=False)  (codec.py:331)
      add_layer(n, 0, "features.conv0", "", -1, cur, 0, c.in_channels, C, 3, 1, 1, 0, c.imsize, c.imsize);
    else
      add_layer(n, 0, "features.In_conv", "", -1, cur, 0, c.in_channels, C, 7, 2, pad0, 0, c.imsize, c.imsize);
    char nm[128];
    for (int bi = 0; bi < c.n_blocks; ++bi) {
      const bool enc = bi < n_enc;
      const int idx = enc ? bi + 1 : bi - n_enc + 1;
      for (int j = 0; j < c.blocks[bi]; ++j) {
        snprintf(nm, sizeof(nm), "features.%sBlock%d.denselayer%d", enc ? "Enc" : "Dec", idx, j + 1);
        const int cin_j = C + j * c.growth_rate, mid_c = c.bottleneck * c.growth_rate;
        if (c.bottleneck > 0 && cin_j > mid_c) {
          // bottleneck form (codec.py:56-64): 1x1 to bn_size * growth channels in a buffer of its own, then the 3x3
          // convolution writes the layer's slice of the block buffer; nn.Dropout2d follows conv2 only (codec.py:70-71)
          const int tmp = add_buf(n, H, H, mid_c);
          add_layer(n, 2, std::string(nm) + ".conv1", std::string(nm) + ".norm1", cur, tmp, 0, cin_j, mid_c, 1, 1, 0, 0, H, H);
          n->layers.back().no_drop = true;
          add_layer(n, 2, std::string(nm) + ".conv2", std::string(nm) + ".norm2", tmp, cur, cin_j, mid_c, c.growth_rate, 3, 1,
                    1, 0, H, H);
          continue;
        }
        add_layer(n, 1, std::string(nm) + ".conv1", std::string(nm) + ".norm1", cur, cur,
                  C + j * c.growth_rate, C + j * c.growth_rate, c.growth_rate, 3, 1, 1, 0, H, H);
      }
      C += c.blocks[bi] * c.growth_rate;
      const bool last = (bi == c.n_blocks - 1);
      if (!last) {
        snprintf(nm, sizeof(nm), "features.Trans%s%d", enc ? "Down" : "Up", idx);
        const int mid = add_buf(n, H, H, C / 2);
        add_layer(n, 2, std::string(nm) + ".conv1", std::string(nm) + ".norm1", cur, mid, 0, C, C / 2, 1,
                  1, 0, 0, H, H);
        const int Hn = enc ? conv_out(H, 3, 2, 1) : 2 * H;
        const int nxt = add_buf(n, Hn, Hn, block_channels(C / 2, c.blocks[bi + 1]));
        // upsample=None: nn.ConvTranspose2d named convT2 (codec.py:139-142)
        const bool convT = !enc && c.upsample == 2;
        add_layer(n, 2, std::string(nm) + (convT ? ".convT2" : ".conv2"), std::string(nm) + ".norm2", mid, nxt, 0,
                  C / 2, C / 2, 3, enc ? 2 : 1, 1, enc ? 0 : 1, H, H);
        n->layers.back().convT = convT;
        C /= 2;
        H = Hn;
        cur = nxt;
      } else {
        const std::string t = "features.LastTransUp";
        const int b1 = add_buf(n, H, H, C / 2);
        add_layer(n, 2, t + ".conv1", t + ".norm1", cur, b1, 0, C, C / 2, 3, 1, 1, 0, H, H);
        // upsample=None: last_decoding adds NO upsampling module (codec.py:176-179): the output stays at H
        const int upl = c.upsample == 2 ? 0 : 1, Ho = upl ? 2 * H : H;
        const int b2 = add_buf(n, Ho, Ho, C / 4);
        add_layer(n, 2, t + ".conv2", t + ".norm2", b1, b2, 0, C / 2, C / 4, 3, 1, 1, upl, H, H);
        add_layer(n, 2, t + ".conv3", t + ".norm3", b2, -1, 0, C / 4, c.out_channels, 5, 1, 2, 0, Ho, Ho);
        n->out_hw = Ho;
        PDES_REQUIRE(decoder || c.upsample == 2 || 2 * H == c.imsize, PDES_ERR_UNSUPPORTED,
                     "imsize %d does not map back to itself through the encoder-decoder (got %d)",
                     c.imsize, 2 * H);
      }
    }
}
  // nn.Dropout2d sites (only live when cfg.dropout): behind every dense-layer convolution (codec.py:70-71),
  // behind conv1 / conv2 of the transitions (110-149) and behind conv1 of the last decoding (171-172)
  {
    int prefix = 0;
    for (auto& L : n->layers) {
      const bool last2 = L.conv_name == "features.LastTransUp.conv2" || L.conv_name == "features.LastTransUp.conv3";
      L.drop = c.dropout != 0 && L.kind != 0 && !last2 && !L.no_drop;
      L.drop_cprefix = prefix;
      if (L.drop) prefix += L.Cout;
    }
  }
  PDES_REQUIRE(c.upsample >= 0 && c.upsample <= 2, PDES_ERR_UNSUPPORTED,
               "upsample mode %d (0 nearest, 1 bilinear, 2 none = transposed convolutions)", c.upsample);
  for (auto& L : n->layers) L.bilinear = L.up && (c.upsample == 1 || L.convT);
  // last consumer of every buffer (first dgrad to run in reverse order stores, the rest accumulate)
  for (size_t b = 0; b < n->bufs.size(); ++b) {
    int lastL = -1;
    for (size_t l = 0; l < n->layers.size(); ++l)
      if (n->layers[l].in_buf == (int)b) lastL = (int)l;
    if (lastL >= 0) n->layers[lastL].last_consumer = true;
  }
  // parameters in named_parameters() order: per module norm.weight, norm.bias, conv.weight
  int64_t off = 0, roff = 0;
  for (auto& L : n->layers) {
    if (L.kind != 0) {
      ParamInfo pw, pb;
      pw.name = L.bn_name + ".weight";
      pw.offset = off;
      pw.ndim = 1;
      pw.shape[0] = L.Cin;
      pw.shape[1] = pw.shape[2] = pw.shape[3] = 1;
      pw.kind = 1;
      L.g_off = off;
      off += pad4(L.Cin);
      pb = pw;
      pb.name = L.bn_name + ".bias";
      pb.offset = off;
      pb.kind = 2;
      L.b_off = off;
      off += pad4(L.Cin);
      n->params.push_back(pw);
      n->params.push_back(pb);
      L.rm_off = roff;
      roff += pad4(L.Cin);
      L.rv_off = roff;
      roff += pad4(L.Cin);
      n->n_bn++;
      if (L.Cin > n->maxC) n->maxC = L.Cin;
    }
    const bool zeros_head = n->coupling && L.out_buf < 0;
    if (zeros_head) {
      // Conv2dZeros (glow_msc.py:240-255): named_parameters() yields the module's own `scale` (1, C, 1, 1)
      // before its child convolution's weight and bias
      ParamInfo psz;
      psz.name = "reduce.conv_zero.scale";
      psz.offset = off;
      psz.ndim = 4;
      psz.shape[0] = 1;
      psz.shape[1] = L.Cout;
      psz.shape[2] = psz.shape[3] = 1;
      psz.kind = 4;
      n->zs_off = off;
      off += pad4(L.Cout);
      n->params.push_back(psz);
    }
    ParamInfo pc;
    pc.name = L.conv_name + ".weight";
    pc.offset = off;
    pc.ndim = 4;
    pc.shape[0] = L.Cout;
    pc.shape[1] = L.Cin;
    pc.shape[2] = pc.shape[3] = L.KS;
    pc.kind = 0;
    L.w_off = off;
    off += pad4((int64_t)L.Cout * L.Cin * L.KS * L.KS);
    n->params.push_back(pc);
    if (zeros_head) {
      ParamInfo pbz;
      pbz.name = L.conv_name + ".bias";
      pbz.offset = off;
      pbz.ndim = 1;
      pbz.shape[0] = L.Cout;
      pbz.shape[1] = pbz.shape[2] = pbz.shape[3] = 1;
      pbz.kind = 3;
      n->zb_off = off;
      off += pad4(L.Cout);
      n->params.push_back(pbz);
    }
  }
  n->param_floats = off;
  n->running_floats = roff;
  // workspace: floats (activations, gradients, packed weights) | doubles | tables
  const int B = c.max_batch;
  size_t f = 0;
  for (auto& b : n->bufs) {
    const size_t sz = (size_t)B * b.H * b.W * b.ld;
    b.act = f;
    f += pad4((int64_t)sz);
    b.grad = f;
    f += pad4((int64_t)sz);
  }
  for (auto& L : n->layers) {
    const size_t taps = (size_t)L.KS * L.KS;
    L.wf = f;
    f += taps * L.CinP * L.CoP;
    L.wb = f;
    f += taps * L.CoutPb * L.CiPb;
    const int pk = (int)(taps * L.CinP * L.CoP + taps * L.CoutPb * L.CiPb);
    if (pk > n->max_pack) n->max_pack = pk;
  }
  for (auto& L : n->layers) {
    // Tensor-core eligibility.  Stride-2 convolutions are run as stride-1 convolutions whose
    // epilogue keeps the even positions (forward) / on a zero-inserted dY (dgrad, wgrad); the first
    // convolution reads the planar network input, the last one writes the planar output.
    const bool in_ok = L.in_buf < 0 || (n->bufs[L.in_buf].ld % 4 == 0);
    const bool out_ok = L.out_buf < 0 || ((n->bufs[L.out_buf].ld % 4 == 0) && (L.coff % 4 == 0));
    const bool sub_ok = L.stride == 1 || (L.stride == 2 && !L.up);
    const bool aligned = in_ok && out_ok && sub_ok;
    if (L.in_buf < 0 && L.out_buf >= 0 && first_conv_supported(L.Cin, L.Cout, L.KS, L.stride)) {
      L.first_k = true;
      continue;
    }
    if (aligned && tc2_supported(L.KS, 1, L.Cin, L.Nf)) {
      L.tc2_fwd = true;
      tc2_plan(L.KS, L.Cin, L.Nf, &L.p2f, n->lowp);
      L.w2f = f;
      f += pad4((int64_t)((L.p2f.pack_elems + 1) / 2));
      n->n_tc2++;
      if (L.p2f.pack_elems > n->max_tc2_pack) n->max_tc2_pack = L.p2f.pack_elems;
    }
    if (aligned && L.tc2_fwd && L.in_buf >= 0 && L.out_buf >= 0 && n->dense_on &&
        dense_fwd_supported(L.KS, L.stride, L.pad, L.up, L.Cin, L.Cout, n->bufs[L.in_buf].H, n->bufs[L.in_buf].W)) {
      L.dense_fwd = true;
      n->n_tc2--;  // its conv_tc2 forward filter is never used: not packed (bind() skips the entry)
      const size_t pe = dense_pack_elems(L.Cin, 16);
      L.wdn = f;
      f += pad4((int64_t)((pe + 1) / 2));
      n->n_tc2++;
      if (pe > n->max_tc2_pack) n->max_tc2_pack = pe;
    }
    if (L.in_buf < 0) continue;  // the first convolution needs no dgrad; its wgrad stays SIMT
    if (aligned && tc2_supported(L.KS, 1, L.Cout, L.Nb)) {
      L.tc2_bwd = true;
      tc2_plan(L.KS, L.Cout, L.Nb, &L.p2b, n->lowp);
      L.w2b = f;
      f += pad4((int64_t)((L.p2b.pack_elems + 1) / 2));
      n->n_tc2++;
      if (L.p2b.pack_elems > n->max_tc2_pack) n->max_tc2_pack = L.p2b.pack_elems;
    }
    if (aligned && L.tc2_bwd && L.out_buf >= 0 && n->dense_bwd_on && wgrad_tc_supported(L.KS, 1) &&
        dense_bwd_supported(L.KS, L.stride, L.pad, L.up, L.Cin, L.Cout, n->bufs[L.in_buf].H, n->bufs[L.in_buf].W)) {
      L.dense_bwd = true;
      n->n_tc2--;  // the conv_tc2 dgrad filter is not packed (bind() skips the entry)
      const size_t pe = dense_bwd_pack_elems(L.Nb);
      L.wdb = f;
      f += pad4((int64_t)((pe + 1) / 2));
      n->n_tc2++;
      if (pe > n->max_tc2_pack) n->max_tc2_pack = pe;
    }
    if (aligned && wgrad_tc_supported(L.KS, 1)) {
      L.tc_wg = true;
      wgrad_tc_dims(L.Cin, L.Cout, &L.ci_pad, &L.co_pad);
      L.dwp = f;
      f += (size_t)L.KS * L.KS * L.ci_pad * L.co_pad;
      n->n_wg++;
      if (L.Cin > n->max_wg_cin) n->max_wg_cin = L.Cin;
      if (L.Cout * 8 * L.KS * L.KS > n->max_wg_elems) n->max_wg_elems = L.Cout * 8 * L.KS * L.KS;
    }
  }
  {
    for (auto& L : n->layers) {
      if (!(L.tc_wg || L.tc2_fwd || L.tc2_bwd)) continue;
      const int Hs = L.in_buf >= 0 ? n->bufs[L.in_buf].H : L.Hs, Ws = L.in_buf >= 0 ? n->bufs[L.in_buf].W : L.Ws;
      const int Hv = L.up ? 2 * Hs : Hs, Wv = L.up ? 2 * Ws : Ws;
      if (L.tc_wg || L.tc2_fwd) {
        L.planes = f;
        f += pad4((int64_t)((act_planes_bytes(B, Hv, Wv, L.Cin) + 3) / 4)) + 64;
      }
      const int zi = L.stride == 2 ? 2 : 1;  // dY planes are zero-inserted for stride-2 layers
      if (L.tc_wg && L.out_buf < 0 && L.stride == 1 && !L.up && L.KS > 1 && L.Cout <= 8 &&
          L.KS * L.KS * L.Cout <= 128) {
        // the last convolution: planar dY with a handful of channels
        L.wg_taps_n = true;
        const int Np = (L.KS * L.KS * L.Cout + 7) & ~7;
        L.planesI = f;
        f += pad4((int64_t)((act_planes_bytes(B, L.Ho, L.Wo, Np) + 3) / 4)) + 64;
        // The same expanded planes are the GEMM-K operand of the layer's data gradient:
        //   dX[q][ci] = sum_(tap, co) I[q][(tap, co)] * W[co][ci][tap]   - one 1x1 GEMM (K = Np) instead of KS^2 taps
        // of K = 16 with 3 useful channels each.  Replaces the conv_tc2 dgrad filter entry of the pack table.
        static const bool on = []() { const char* e = getenv("PDES_DGRAD_IM2COL"); return !(e && e[0] == '0'); }();
        if (on && L.in_buf >= 0 && L.tc2_bwd && !L.dense_bwd && tc2_supported(1, 1, Np, L.Nb)) {
          L.dg_im2col = true;
          tc2_plan(1, Np, L.Nb, &L.p2i, n->lowp);
          L.w2i = f;
          f += pad4((int64_t)((L.p2i.pack_elems + 1) / 2));
          if (L.p2i.pack_elems > n->max_tc2_pack) n->max_tc2_pack = L.p2i.pack_elems;
        }
      }
      if (L.in_buf >= 0 && (L.tc_wg || L.tc2_bwd)) {
        L.planesB = f;
        f += pad4((int64_t)((act_planes_bytes(B, zi * L.Ho, zi * L.Wo, L.Cout) + 3) / 4)) + 64;
      }
    }
  }
  {
    size_t mx = 0;
    for (const auto& L : n->layers)
      if (L.bilinear) {
        const Buf& ib = n->bufs[L.in_buf];
        const size_t sz = (size_t)B * 4 * ib.H * ib.W * rup(L.Cin, 4);
        if (sz > mx) mx = sz;
      }
    n->up_scratch = f;
    f += pad4((int64_t)mx);
    for (auto& L : n->layers)
      if (L.convT) {
        L.wt = f;
        f += pad4((int64_t)L.Cout * L.Cin * L.KS * L.KS);
        L.gt = f;
        f += pad4((int64_t)L.Cout * L.Cin * L.KS * L.KS);
      }
  }
  if (n->coupling) {
    n->out_keep = f;
    f += pad4((int64_t)B * c.out_channels * n->out_hw * n->out_hw);
    n->dyg = f;
    f += pad4((int64_t)B * c.out_channels * n->out_hw * n->out_hw);
  }
  n->dyinv = f;
  f += pad4((int64_t)n->layers.size());
  n->xin = f;
  f += pad4((int64_t)B * c.in_channels * n->in_hw * n->in_hw);
  n->xs = f;
  f += pad4((int64_t)B * c.in_channels * n->in_hw * n->in_hw);
  n->outs = f;
  f += pad4((int64_t)B * c.out_channels * n->out_hw * n->out_hw);
  n->douts = f;
  f += pad4((int64_t)B * c.out_channels * n->out_hw * n->out_hw);
  n->ws_floats = f;
  size_t d = 0;
  for (auto& b : n->bufs) {
    b.stat = d;
    d += 2 * (size_t)b.C;
  }
  for (auto& L : n->layers) {
    L.bsum = d;
    d += 2 * (size_t)L.Cin;
  }
  n->gmax_off = d;
  d += (n->bufs.size() + 1 + 1) / 2;  // two unsigned words per double slot
  n->ws_doubles = d;
  n->off_doubles = (n->ws_floats * sizeof(float) + 255) & ~(size_t)255;
  n->off_tables = (n->off_doubles + n->ws_doubles * sizeof(double) + 255) & ~(size_t)255;
  n->ws_bytes = n->off_tables + sizeof(PackDesc) * n->layers.size() +
                sizeof(BnLayerDesc) * (size_t)n->n_bn +
                sizeof(TcWgradUnpack) * (size_t)n->n_wg + sizeof(Tc2PackDesc) * (size_t)n->n_tc2 + 256;
  return PDES_OK;
}

inline float* wsf(const pdes_net* n, size_t off) { return reinterpret_cast<float*>(n->ws) + off; }
inline double* wsd(const pdes_net* n, size_t off) {
  return reinterpret_cast<double*>(n->ws + n->off_doubles) + off;
}
inline unsigned* gmax_slot(const pdes_net* n, int buf) {  // buf < 0: the network output gradient
  return reinterpret_cast<unsigned*>(wsd(n, n->gmax_off)) + (buf < 0 ? (int)n->bufs.size() : buf);
}
inline PackDesc* pack_table(const pdes_net* n) { return reinterpret_cast<PackDesc*>(n->ws + n->off_tables); }
inline BnLayerDesc* bn_table(const pdes_net* n) {
  return reinterpret_cast<BnLayerDesc*>(n->ws + n->off_tables + sizeof(PackDesc) * n->layers.size());
}

inline TcWgradUnpack* wg_table(const pdes_net* n) {
  return reinterpret_cast<TcWgradUnpack*>(n->ws + n->off_tables + sizeof(PackDesc) * n->layers.size() +
                                          sizeof(BnLayerDesc) * (size_t)n->n_bn);
}

inline Tc2PackDesc* tc2_table(const pdes_net* n) {
  return reinterpret_cast<Tc2PackDesc*>(reinterpret_cast<unsigned char*>(wg_table(n)) +
                                        sizeof(TcWgradUnpack) * (size_t)n->n_wg);
}

BnSrc bn_src(const pdes_net* n, const Layer& L, int B, bool training) {
  BnSrc s;
  memset(&s, 0, sizeof(s));
  const Buf& ib = n->bufs[L.in_buf];
  s.gamma = n->p + L.g_off;
  s.beta = n->p + L.b_off;
  s.eps = 1e-5f;
  if (training) {
    s.sum = wsd(n, ib.stat);
    s.sumsq = wsd(n, ib.stat) + ib.C;
    s.inv_count = 1.0 / ((double)B * ib.H * ib.W);
  } else {
    s.use_running = 1;
    s.run_mean = n->run + L.rm_off;
    s.run_var = n->run + L.rv_off;
  }
  return s;
}

}  // namespace

extern "C" int pdes_densenet_create(const pdes_densenet_config* cfg, pdes_net_t** out) {
  PDES_REQUIRE(cfg && out, PDES_ERR_INVALID, "pdes_densenet_create: null argument");
  pdes_net* n = new pdes_net();
  n->cfg = *cfg;
  {
    const char* e = getenv("PDES_DENSE_FWD");
    n->dense_on = (e && e[0] == '0') ? 0 : 1;
    const char* e2 = getenv("PDES_DENSE_BWD");
    n->dense_bwd_on = (e2 && e2[0] == '0') ? 0 : 1;
    if (cfg->dropout) n->dense_bwd_on = 0;  // the fused dgrad folds the dY correction; dropout needs it standalone
    const char* e3 = getenv("PDES_CONV_DTYPE");
    n->lowp = (e3 && !strcmp(e3, "bf16")) ? LOWP_BF16 : ((e3 && !strcmp(e3, "fp16")) ? LOWP_FP16 : LOWP_NONE);
  }
  const int rc = build(n);
  if (rc != PDES_OK) {
    delete n;
    return rc;
  }
  {
    // PDES_WGRAD_STREAMS = 0 | 1 | 2 side streams for the weight-gradient kernels (default 1;
    // measured on B200: eager step 4.36 -> 3.97 ms with one, 4.07 ms with two).  The streams and
    // events themselves are created by bind(), on the device that owns the workspace: this
    // constructor touches no device (it also serves as a pure layout query).
    const char* e = getenv("PDES_WGRAD_STREAMS");
    n->n_side_req = e ? atoi(e) : 1;
    if (n->n_side_req < 0) n->n_side_req = 0;
    if (n->n_side_req > 2) n->n_side_req = 2;
    n->n_side = 0;
  }
  *out = n;
  return PDES_OK;
}

namespace {
void drop_side_streams(pdes_net* n) {
  for (int k = 0; k < 2; ++k) {
    if (n->side[k]) cudaStreamDestroy(n->side[k]);
    if (n->ev_fork[k]) cudaEventDestroy(n->ev_fork[k]);
    if (n->ev_join[k]) cudaEventDestroy(n->ev_join[k]);
    n->side[k] = nullptr;
    n->ev_fork[k] = n->ev_join[k] = nullptr;
  }
  n->n_side = 0;
  n->side_dev = -1;
}
// side streams / fork-join events live on the device current at bind() time
void ensure_side_streams(pdes_net* n) {
  const int dev = cur_device();
  if (n->side_dev == dev) return;
  drop_side_streams(n);
  n->side_dev = dev;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);  // lo = least priority
  for (int k = 0; k < n->n_side_req; ++k) {
    if (cudaStreamCreateWithPriority(&n->side[k], cudaStreamNonBlocking, lo) != cudaSuccess ||
        cudaEventCreateWithFlags(&n->ev_fork[k], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&n->ev_join[k], cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      drop_side_streams(n);
      n->side_dev = dev;
      return;
    }
    n->n_side = k + 1;
  }
}
}  // namespace

// The executor graphs bake in the launch structure: anything that changes it (implementation switches,
// timing marks, re-binding) drops them; the next calls capture again.
static void drop_exec_graphs(pdes_net_t* n) {
  for (auto& gs : n->gfwd) {
    if (gs.exec) cudaGraphExecDestroy(gs.exec);
    gs = pdes_net::GraphSlot();
  }
  for (auto& gs : n->gbwd) {
    if (gs.exec) cudaGraphExecDestroy(gs.exec);
    gs = pdes_net::GraphSlot();
  }
}

extern "C" void pdes_densenet_destroy(pdes_net_t* net) {
  if (!net) return;
  drop_exec_graphs(net);
  if (net->cap_stream) cudaStreamDestroy(net->cap_stream);
  drop_side_streams(net);
  delete net;
}

extern "C" int pdes_densenet_num_params(const pdes_net_t* n) { return n ? (int)n->params.size() : 0; }
extern "C" int64_t pdes_densenet_param_floats(const pdes_net_t* n) { return n ? n->param_floats : 0; }
extern "C" int pdes_densenet_param_info(const pdes_net_t* n, int idx, char* name, size_t cap,
                                        int64_t* offset, int32_t* ndim, int64_t shape[4],
                                        int32_t* kind) {
  PDES_REQUIRE(n && idx >= 0 && idx < (int)n->params.size(), PDES_ERR_INVALID,
               "pdes_densenet_param_info: index %d out of range", idx);
  const ParamInfo& p = n->params[idx];
  if (name && cap) {
    strncpy(name, p.name.c_str(), cap - 1);
    name[cap - 1] = 0;
  }
  if (offset) *offset = p.offset;
  if (ndim) *ndim = p.ndim;
  if (shape)
    for (int i = 0; i < 4; ++i) shape[i] = p.shape[i];
  if (kind) *kind = p.kind;
  return PDES_OK;
}
extern "C" int pdes_densenet_num_bn(const pdes_net_t* n) { return n ? n->n_bn : 0; }
extern "C" int64_t pdes_densenet_running_floats(const pdes_net_t* n) { return n ? n->running_floats : 0; }
extern "C" int pdes_densenet_bn_info(const pdes_net_t* n, int idx, char* name, size_t cap,
                                     int64_t* mean_offset, int64_t* var_offset, int32_t* channels) {
  PDES_REQUIRE(n && idx >= 0 && idx < n->n_bn, PDES_ERR_INVALID,
               "pdes_densenet_bn_info: index %d out of range", idx);
  int k = -1;
  for (const auto& L : n->layers) {
    if (L.kind == 0) continue;
    if (++k == idx) {
      if (name && cap) {
        strncpy(name, L.bn_name.c_str(), cap - 1);
        name[cap - 1] = 0;
      }
      if (mean_offset) *mean_offset = L.rm_off;
      if (var_offset) *var_offset = L.rv_off;
      if (channels) *channels = L.Cin;
      return PDES_OK;
    }
  }
  return PDES_ERR_INVALID;
}

extern "C" int pdes_densenet_output_size(const pdes_net_t* n) { return n ? n->out_hw : 0; }

extern "C" int pdes_densenet_dropout_sites(const pdes_net_t* n, int32_t* channels, int cap) {
  if (!n) return 0;
  int k = 0;
  for (const auto& L : n->layers) {
    if (!L.drop) continue;
    if (channels && k < cap) channels[k] = L.Cout;
    ++k;
  }
  return k;
}

extern "C" int pdes_densenet_set_dropout(pdes_net_t* n, const float* masks) {
  PDES_REQUIRE(n, PDES_ERR_INVALID, "pdes_densenet_set_dropout: null net");
  PDES_REQUIRE(masks == nullptr || n->cfg.dropout, PDES_ERR_STATE,
               "pdes_densenet_set_dropout: the network was created without dropout");
  n->drop_masks = masks;
  return PDES_OK;
}

extern "C" size_t pdes_densenet_workspace_bytes(const pdes_net_t* n) { return n ? n->ws_bytes : 0; }

// filter / filter-gradient pointers the convolution kernels see: the parameter itself, or (transposed
// convolution) the equivalent Conv2d filter in the workspace and its self-cleaning gradient staging buffer
static inline const float* conv_w(const pdes_net_t* n, const Layer& L) {
  return L.convT ? wsf(const_cast<pdes_net_t*>(n), L.wt) : n->p + L.w_off;
}
static inline float* conv_dw(pdes_net_t* n, const Layer& L) { return L.convT ? wsf(n, L.gt) : n->g + L.w_off; }

extern "C" int pdes_densenet_bind(pdes_net_t* n, float* params, float* grads, float* running,
                                  void* workspace, size_t workspace_bytes) {
  PDES_REQUIRE(n && params && running && workspace, PDES_ERR_INVALID, "pdes_densenet_bind: null pointer");
  PDES_REQUIRE(workspace_bytes >= n->ws_bytes, PDES_ERR_INVALID,
               "pdes_densenet_bind: workspace %zu < %zu bytes", workspace_bytes, n->ws_bytes);
  PDES_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)running | (uintptr_t)workspace) & 15u) == 0,
               PDES_ERR_INVALID, "pdes_densenet_bind: buffers must be 16-byte aligned");
  // The caller zero-fills the workspace on ITS stream (possibly a non-blocking one that the
  // synchronous table uploads below do not order against): wait for everything first.
  PDES_CUDA(cudaDeviceSynchronize());
  ensure_side_streams(n);
  drop_exec_graphs(n);
  {
    // Executor graphs (the module API: forward / backward called from a Python loop): from the second call with
    // the same batch size on, the launches of a pass are replayed as ONE CUDA graph captured on a private
    // stream (the caller's stream may be the legacy default stream, which cannot be captured) and launched into
    // the caller's stream.  PDES_EXEC_GRAPH=0 keeps direct launches.  (Round 1 tried to capture on the caller's
    // stream, which silently failed on torch's default stream: "no gain" was the fallback path.)
    const char* e = getenv("PDES_EXEC_GRAPH");
    n->use_graph = (e && e[0] == '0') ? 0 : 1;
    if (n->use_graph && !n->cap_stream &&
        cudaStreamCreateWithFlags(&n->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
      cudaGetLastError();
      n->cap_stream = nullptr;
      n->use_graph = 0;
    }
  }
  n->p = params;
  n->g = grads;
  n->run = running;
  n->ws = (unsigned char*)workspace;
  n->fwd_train_done = false;
  std::vector<PackDesc> pt(n->layers.size());
  std::vector<BnLayerDesc> bt;
  for (size_t i = 0; i < n->layers.size(); ++i) {
    const Layer& L = n->layers[i];
    PackDesc& d = pt[i];
    d.w = conv_w(n, L);
    d.wf = wsf(n, L.wf);
    d.wb = L.in_buf >= 0 ? wsf(n, L.wb) : nullptr;
    d.Cout = L.Cout;
    d.Cin = L.Cin;
    d.KS = L.KS;
    d.CinP = L.CinP;
    d.CoP = L.CoP;
    d.CoutPb = L.CoutPb;
    d.CiPb = L.CiPb;
    if (L.kind != 0) {
      const Buf& ib = n->bufs[L.in_buf];
      BnLayerDesc b;
      b.sum = wsd(n, ib.stat);
      b.sumsq = wsd(n, ib.stat) + ib.C;
      b.bsum = wsd(n, L.bsum);
      b.run_mean = n->run + L.rm_off;
      b.run_var = n->run + L.rv_off;
      b.dgamma = n->g ? n->g + L.g_off : nullptr;
      b.dbeta = n->g ? n->g + L.b_off : nullptr;
      b.count = (double)ib.H * ib.W;  // per sample; multiplied by B on the device
      b.C = L.Cin;
      bt.push_back(b);
    }
  }
  std::vector<Tc2PackDesc> t2;
  for (const auto& L : n->layers) {
    for (int dir = 0; dir < 2; ++dir) {
      if (!(dir == 0 ? L.tc2_fwd : L.tc2_bwd)) continue;
      if (dir == 0 && L.dense_fwd) continue;
      if (dir == 1 && L.dense_bwd) continue;
      if (dir == 1 && L.dg_im2col) {
        Tc2PackDesc d;
        d.w = conv_w(n, L);
        d.dst = reinterpret_cast<op16*>(wsf(n, L.w2i));
        d.Cout = L.Cout;
        d.Cin = L.Cin;
        d.KS = L.KS;
        d.N = L.Nb;
        d.KC = L.p2i.KC;
        d.nchunks = L.p2i.nchunks;
        d.transpose = 0;
        d.dxn = 3;
        d.CoP = 0;
        d.lowp = n->lowp;
        t2.push_back(d);
        continue;
      }
      Tc2PackDesc d;
      d.w = conv_w(n, L);
      d.dst = reinterpret_cast<op16*>(wsf(n, dir == 0 ? L.w2f : L.w2b));
      d.Cout = L.Cout;
      d.Cin = L.Cin;
      d.KS = L.KS;
      d.N = dir == 0 ? L.Nf : L.Nb;
      d.KC = dir == 0 ? L.p2f.KC : L.p2b.KC;
      d.nchunks = dir == 0 ? L.p2f.nchunks : L.p2b.nchunks;
      d.transpose = dir;
      d.dxn = 0;
      d.CoP = 0;
      d.lowp = n->lowp;
      t2.push_back(d);
    }
    if (L.dense_fwd) {
      Tc2PackDesc d;
      d.w = conv_w(n, L);
      d.dst = reinterpret_cast<op16*>(wsf(n, L.wdn));
      d.Cout = L.Cout;
      d.Cin = L.Cin;
      d.KS = L.KS;
      d.N = 3 * 16;
      d.KC = 32;
      d.nchunks = (L.Cin + 31) / 32;
      d.transpose = 0;
      d.dxn = 1;
      d.CoP = 16;
      d.lowp = n->lowp;
      t2.push_back(d);
    }
    if (L.dense_bwd) {
      Tc2PackDesc d;
      d.w = conv_w(n, L);
      d.dst = reinterpret_cast<op16*>(wsf(n, L.wdb));
      d.Cout = L.Cout;
      d.Cin = L.Cin;
      d.KS = L.KS;
      d.N = L.Nb;
      d.KC = 48;
      d.nchunks = 1;
      d.transpose = 0;
      d.dxn = 2;
      d.CoP = 16;
      d.lowp = n->lowp;
      t2.push_back(d);
    }
  }
  if (!t2.empty())
    PDES_CUDA(cudaMemcpy(tc2_table(n), t2.data(), sizeof(Tc2PackDesc) * t2.size(), cudaMemcpyHostToDevice));
  std::vector<TcWgradUnpack> wt;
  for (const auto& L : n->layers) {
    if (!L.tc_wg || !n->g) continue;
    TcWgradUnpack u;
    memset(&u, 0, sizeof(u));
    u.dw = conv_dw(n, L);
    u.dwp = wsf(n, L.dwp);
    u.Cout = L.Cout;
    u.Cin = L.Cin;
    u.KS = L.KS;
    u.ci_pad = L.ci_pad;
    u.co_pad = L.co_pad;
    u.taps_in_n = 0;
    if (L.wg_taps_n) {
      u.taps_in_n = L.KS * L.KS;
      u.co_pad = (L.KS * L.KS * L.Cout + 15) / 16 * 16;
    }
    wt.push_back(u);
  }
  n->n_wg_bound = (int)wt.size();
  if (!wt.empty())
    PDES_CUDA(cudaMemcpy(wg_table(n), wt.data(), sizeof(TcWgradUnpack) * wt.size(), cudaMemcpyHostToDevice));
  PDES_CUDA(cudaMemcpy(pack_table(n), pt.data(), sizeof(PackDesc) * pt.size(), cudaMemcpyHostToDevice));
  if (!bt.empty())
    PDES_CUDA(cudaMemcpy(bn_table(n), bt.data(), sizeof(BnLayerDesc) * bt.size(), cudaMemcpyHostToDevice));
  return PDES_OK;
}

extern "C" int pdes_densenet_set_conv_impl(pdes_net_t* n, int impl) {
  // 0 = tcgen05 (two-piece fp16) where supported, 1 = CUDA-core fp32 everywhere, 2 = same as 0,
  // 3 / 4 / 5 = tcgen05 for the forward only / the dgrad only / the wgrad only (diagnostics),
  // 6 = the launch structure of 3..5 (dedicated first-convolution kernels) with NO tensor-core kernel:
  // the exact-fp32 baseline whose forward is bitwise the forward of 4 and 5
  // 7 = exact-fp32 forward of 6 with BOTH backward kernels on tcgen05 (the fused thin-layer dgrad needs both)
  PDES_REQUIRE(n && impl >= 0 && impl <= 7, PDES_ERR_INVALID, "pdes_densenet_set_conv_impl: impl in 0..7");
  drop_exec_graphs(n);
  n->conv_impl = impl == 1 ? 1 : 0;
  n->tc_mask = impl == 3 ? 1 : (impl == 4 ? 2 : (impl == 5 ? 4 : (impl == 6 ? 0 : (impl == 7 ? 6 : 7))));
  return PDES_OK;
}

extern "C" int pdes_densenet_set_timing(pdes_net_t* n, int on) {
  PDES_REQUIRE(n, PDES_ERR_INVALID, "pdes_densenet_set_timing: null net");
  for (auto& m : n->marks) cudaEventDestroy(m.second);
  n->marks.clear();
  n->timing = on != 0;
  drop_exec_graphs(n);
  return PDES_OK;
}

// Synchronises the device and prints "<us> <label>" per launch since set_timing(1) to stderr.
extern "C" int pdes_densenet_timing_report(pdes_net_t* n) {
  PDES_REQUIRE(n, PDES_ERR_INVALID, "pdes_densenet_timing_report: null net");
  PDES_CUDA(cudaDeviceSynchronize());
  for (size_t i = 1; i < n->marks.size(); ++i) {
    if (n->marks[i].first[0] == '@') continue;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, n->marks[i - 1].second, n->marks[i].second);
    fprintf(stderr, "[pdes timing] %9.2f us  %s\n", ms * 1e3f, n->marks[i].first.c_str());
  }
  for (auto& m : n->marks) cudaEventDestroy(m.second);
  n->marks.clear();
  return PDES_OK;
}

// Same measurements as pdes_densenet_timing_report, returned to the caller: us[i] / flops[i] of launch i
// (flops = useful 2*MAC of a convolution launch, 0 otherwise) and the '\n'-separated labels.
// Returns the number of launches (<= cap filled) or a negative error code; clears the marks.
extern "C" int pdes_densenet_timing_read(pdes_net_t* n, float* us, double* flops, int cap, char* labels,
                                         size_t labels_cap) {
  if (!n || !us || !flops || !labels || labels_cap == 0) return -PDES_ERR_INVALID;
  if (cudaDeviceSynchronize() != cudaSuccess) return -PDES_ERR_CUDA;
  int k = 0;
  size_t pos = 0;
  labels[0] = 0;
  for (size_t i = 1; i < n->marks.size(); ++i) {
    if (n->marks[i].first[0] == '@') continue;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, n->marks[i - 1].second, n->marks[i].second);
    if (k < cap) {
      us[k] = ms * 1e3f;
      flops[k] = n->marks[i].flops;
      const std::string& lb = n->marks[i].first;
      if (pos + lb.size() + 2 < labels_cap) {
        memcpy(labels + pos, lb.c_str(), lb.size());
        pos += lb.size();
        labels[pos++] = '\n';
        labels[pos] = 0;
      }
    }
    ++k;
  }
  for (auto& m : n->marks) cudaEventDestroy(m.second);
  n->marks.clear();
  return k;
}

extern "C" int pdes_densenet_last_launches(const pdes_net_t* n) { return n ? n->launches : 0; }

extern "C" double pdes_densenet_flops(const pdes_net_t* n, int B, int training) {
  if (!n) return 0.0;
  double f = 0.0;
  for (const auto& L : n->layers) {
    const double one = 2.0 * L.Cin * L.Cout * L.KS * L.KS * (double)L.Ho * L.Wo * B;
    f += one;
    if (training) f += one + (L.in_buf >= 0 ? one : 0.0);
  }
  return f;
}

static int forward_impl(pdes_net_t* n, const float* x, float* out, int B, int training, void* stream) {
  PDES_REQUIRE(n && n->ws, PDES_ERR_STATE, "pdes_densenet_forward: bind() first");
  PDES_REQUIRE(x && out, PDES_ERR_INVALID, "pdes_densenet_forward: null pointer");
  PDES_REQUIRE(B >= 1 && B <= n->cfg.max_batch, PDES_ERR_INVALID,
               "pdes_densenet_forward: batch %d outside 1..%d", B, n->cfg.max_batch);
  cudaStream_t st = (cudaStream_t)stream;
  n->launches = 0;
  const bool tr = training != 0;
  mark(n, st, "@forward");
  if (tr) {
    PDES_CUDA(cudaMemsetAsync(wsd(n, 0), 0, n->ws_doubles * sizeof(double), st));
    n->launches++;
    mark(n, st, "memset stats");
  }
  int rc = PDES_OK;
  for (const auto& L : n->layers)
    if (L.convT) {
      rc = launch_convt_weight(n->p + L.w_off, wsf(n, L.wt), L.Cin, L.Cout, L.KS, st);
      if (rc) return rc;
      n->launches++;
      mark(n, st, "convT_weight " + L.conv_name);
    }
  bool need_simt_pack = n->conv_impl != 0 || n->tc_mask != 7;
  for (const auto& L : n->layers)
    if (!(L.tc2_fwd || L.first_k) || (L.in_buf >= 0 && !L.tc2_bwd)) need_simt_pack = true;
  if (need_simt_pack) {
    rc = launch_pack_weights(pack_table(n), (int)n->layers.size(), n->max_pack, st);
    if (rc) return rc;
    n->launches++;
    mark(n, st, "pack_weights");
  }
  // The filter packing is independent of the first convolution (dedicated CUDA-core kernel on the raw
  // filter): it runs beside it on a side stream and joins before the first tensor-core layer.
  bool pack_forked = false;
  if (n->conv_impl == 0 && n->n_tc2 > 0) {
    cudaStream_t pst = st;
    if (n->n_side > 0 && n->side[0] && !n->timing && !n->layers.empty() && n->layers[0].first_k && !n->coupling) {
      PDES_CUDA(cudaEventRecord(n->ev_fork[0], st));
      PDES_CUDA(cudaStreamWaitEvent(n->side[0], n->ev_fork[0], 0));
      pst = n->side[0];
      pack_forked = true;
    }
    rc = launch_pack_tc2(tc2_table(n), n->n_tc2, n->max_tc2_pack, pst);
    if (rc) return rc;
    n->launches++;
    mark(n, st, "pack_tc2");
    if (pack_forked) PDES_CUDA(cudaEventRecord(n->ev_join[0], n->side[0]));
  }
  if (n->coupling) {
    const Buf& b0 = n->bufs[0];
    rc = launch_nchw_to_block(x, wsf(n, b0.act), b0.ld, n->cfg.in_channels, B, b0.H * b0.W,
                              tr ? wsd(n, b0.stat) : nullptr, tr ? wsd(n, b0.stat) + b0.C : nullptr, st);
    if (rc) return rc;
    n->launches++;
    mark(n, st, "input->block");
  } else if (tr) {
    PDES_CUDA(cudaMemcpyAsync(wsf(n, n->xin), x,
                              sizeof(float) * (size_t)B * n->cfg.in_channels * n->in_hw * n->in_hw,
                              cudaMemcpyDeviceToDevice, st));
    n->launches++;
    mark(n, st, "copy xin");
  }
  for (const auto& L : n->layers) {
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    if (L.in_buf < 0) {
      a.x = tr ? wsf(n, n->xin) : x;
      a.in_nchw = 1;
      a.Cin = L.Cin;
      a.Hs = L.Hs;
      a.Ws = L.Ws;
    } else {
      const Buf& ib = n->bufs[L.in_buf];
      a.x = wsf(n, ib.act);
      a.ldx = ib.ld;
      a.Cin = L.Cin;
      a.Hs = ib.H;
      a.Ws = ib.W;
      a.pro = 1;
      a.bn = bn_src(n, L, B, tr);
    }
    a.B = B;
    a.in_mode = L.up ? IN_UPSAMPLE : IN_DIRECT;
    a.w = wsf(n, L.wf);
    a.CinP = L.CinP;
    a.CoP = L.CoP;
    a.Cout = L.Cout;
    a.KS = L.KS;
    a.pad = L.pad;
    a.stride = L.stride;
    a.Ho = L.Ho;
    a.Wo = L.Wo;
    if (L.out_buf < 0) {
      a.epi = EPI_NCHW;
      a.y = out;
    } else {
      const Buf& ob = n->bufs[L.out_buf];
      a.epi = EPI_NHWC;
      a.y = wsf(n, ob.act);
      a.ldy = ob.ld;
      a.coff = L.coff;
      if (tr) {
        a.o_sum = wsd(n, ob.stat) + L.coff;
        a.o_sumsq = wsd(n, ob.stat) + ob.C + L.coff;
      }
    }
    if (L.bilinear) {
      // a_up = bilinear(relu(bn(x))) once, fp32 NHWC; the convolution below then sees a plain direct input
      const Buf& ib = n->bufs[L.in_buf];
      BilinearArgs ba;
      memset(&ba, 0, sizeof(ba));
        ba.zero_insert = L.convT ? 1 : 0;
      ba.x = a.x;
      ba.ldx = a.ldx;
      ba.C = L.Cin;
      ba.H = ib.H;
      ba.W = ib.W;
      ba.B = B;
      ba.pro = a.pro;
      ba.bn = a.bn;
      ba.up = wsf(n, n->up_scratch);
      ba.ldu = rup(L.Cin, 4);
      rc = launch_bilinear_up(ba, st);
      if (rc) return rc;
      n->launches++;
      mark(n, st, "bilinear.f " + L.conv_name);
      a.x = ba.up;
      a.ldx = ba.ldu;
      a.Hs = 2 * ib.H;
      a.Ws = 2 * ib.W;
      a.pro = 0;
      a.in_mode = IN_DIRECT;
    }
    const bool dropping = tr && n->drop_masks != nullptr && L.drop && L.out_buf >= 0;
    double* drop_sum = a.o_sum;
    double* drop_sumsq = a.o_sumsq;
    if (dropping) a.o_sum = a.o_sumsq = nullptr;   // the statistics are those of the MASKED output (below)
    const bool use_dense = n->conv_impl == 0 && L.dense_fwd && (n->tc_mask & 1);
    const bool want_planes = n->conv_impl == 0 && (L.tc2_fwd || (tr && L.tc_wg)) && !use_dense;
    const int Hs_l = L.in_buf >= 0 ? n->bufs[L.in_buf].H : L.Hs, Ws_l = L.in_buf >= 0 ? n->bufs[L.in_buf].W : L.Ws;
    if (want_planes) {
      // fp16 pieces of relu(bn(x)) (nearest-upsampled if needed): read by the forward conv and,
      // in training, again by the weight-gradient kernel
      ActSplitArgs sa;
      memset(&sa, 0, sizeof(sa));
      sa.x = a.x;
      sa.ldx = a.ldx;
      sa.nchw = a.in_nchw;
      sa.C = L.Cin;
      sa.Hs = L.bilinear ? 2 * Hs_l : Hs_l;   // (bilinear: a.x is the already upsampled activation)
      sa.Ws = L.bilinear ? 2 * Ws_l : Ws_l;
      sa.B = B;
      sa.up = L.bilinear ? 0 : L.up;
      sa.pro = a.pro;
      sa.bn = a.bn;
      sa.out = reinterpret_cast<op16*>(wsf(n, L.planes));
      sa.Cp = (L.Cin + 7) & ~7;
      sa.scale = pow2f(kActScaleLog2);
      sa.lowp = n->lowp;
      rc = launch_act_split(sa, st);
      if (rc) return rc;
      n->launches++;
      mark(n, st, "split.f " + L.conv_name);
    }
    if (n->conv_impl == 0 && L.first_k) {
      FirstConvArgs fa;
      memset(&fa, 0, sizeof(fa));
      fa.x = a.x;
      fa.w = conv_w(n, L);
      fa.y = a.y;
      fa.ldy = a.ldy;
      fa.coff = a.coff;
      fa.o_sum = a.o_sum;
      fa.o_sumsq = a.o_sumsq;
      fa.B = B;
      fa.Cin = L.Cin;
      fa.H = L.Hs;
      fa.W = L.Ws;
      fa.Cout = L.Cout;
      fa.KS = L.KS;
      fa.stride = L.stride;
      fa.pad = L.pad;
      fa.Ho = L.Ho;
      fa.Wo = L.Wo;
      rc = launch_first_conv_fwd(fa, st);
      if (pack_forked) {   // join: every later layer reads the packed filters
        PDES_CUDA(cudaStreamWaitEvent(st, n->ev_join[0], 0));
        pack_forked = false;
      }
    } else if (use_dense) {
      // thin layer: BatchNorm + ReLU + fp16 split happen inside the convolution kernel, which also emits
      // the operand planes of the weight gradient (training)
      DenseFwdArgs da;
      memset(&da, 0, sizeof(da));
      da.x = a.x;
      da.ldx = a.ldx;
      da.Cin = L.Cin;
      da.H = Hs_l;
      da.W = Ws_l;
      da.B = B;
      da.pro = a.pro;
      da.bn = a.bn;
      da.wpk = reinterpret_cast<const op16*>(wsf(n, L.wdn));
      da.CoP = 16;
      da.Cout = L.Cout;
      da.y = a.y;
      da.ldy = a.ldy;
      da.coff = a.coff;
      da.o_sum = a.o_sum;
      da.o_sumsq = a.o_sumsq;
      da.planes = (tr && L.tc_wg) ? reinterpret_cast<op16*>(wsf(n, L.planes)) : nullptr;
      da.Cp = (L.Cin + 7) & ~7;
      da.out_scale = pow2f(-(kActScaleLog2 + kWScaleLog2));
      da.lowp = n->lowp;
      da.b_early = 1;  // packed by pack_tc2 at the head of the forward pass, >= 2 launches ago
      {
        // PDES_DENSE_DBG=<layer index>: phase timestamps of that layer's kernel, printed after the launch
        static int dbg_layer = -2;
        static long long* dbg_buf = nullptr;
        if (dbg_layer == -2) {
          const char* e = getenv("PDES_DENSE_DBG");
          dbg_layer = e ? atoi(e) : -1;
          if (dbg_layer >= 0 && cudaMalloc((void**)&dbg_buf, sizeof(long long) * 256) != cudaSuccess) dbg_layer = -1;
        }
        {
          static int exp_v = -1;
          if (exp_v < 0) {
            const char* e = getenv("PDES_DENSE_EXP");
            exp_v = e ? atoi(e) : 0;
          }
          da.exp = exp_v;
        }
        const int my_index = (int)(&L - &n->layers[0]);
        if (dbg_layer == my_index && dbg_buf != nullptr && !stream_capturing(st)) {
          cudaMemsetAsync(dbg_buf, 0, sizeof(long long) * 256, st);
          da.dbg = dbg_buf;
          rc = launch_conv_dense_fwd(da, st);
          long long h[256];
          cudaMemcpyAsync(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost, st);
          cudaStreamSynchronize(st);
          for (int c = 0; c < 2; ++c) {
            fprintf(stderr, "[dense dbg] %s CTA %d:", L.conv_name.c_str(), c);
            for (int i = 1; i < 64; ++i)
              if (h[c * 64 + i]) fprintf(stderr, " s%d=+%lld", i, h[c * 64 + i] - h[c * 64]);
            fprintf(stderr, "\n");
          }
        } else {
          rc = launch_conv_dense_fwd(da, st);
        }
      }
    } else if (n->conv_impl == 0 && L.tc2_fwd && (n->tc_mask & 1)) {
      const int Hv = L.up ? 2 * Hs_l : Hs_l, Wv = L.up ? 2 * Ws_l : Ws_l;
      Tc2Args t;
      memset(&t, 0, sizeof(t));
      t.c = a;
      if (L.stride == 2) {  // stride-1 evaluation + subsampled store
        t.c.stride = 1;
        t.c.Ho = Hv + 2 * L.pad - L.KS + 1;
        t.c.Wo = Wv + 2 * L.pad - L.KS + 1;
        t.osub = 1;
      }
      t.wpk = reinterpret_cast<const op16*>(wsf(n, L.w2f));
      t.lowp = n->lowp;
      t.out_scale = pow2f(-(kActScaleLog2 + kWScaleLog2));
      t.N = L.Nf;
      t.KC = L.p2f.KC;
      t.nchunks = L.p2f.nchunks;
      t.ngroups = L.p2f.ngroups;
      t.S = L.p2f.S;
      t.TS = L.p2f.TS;
      t.AST = L.p2f.AST;
      t.NB = L.p2f.NB;
      t.TPB = L.p2f.TPB;
      rc = launch_conv_tc2(t, reinterpret_cast<const op16*>(wsf(n, L.planes)), Hv, Wv, L.Cin, st);
    } else {
      rc = launch_conv_simt(a, st);
    }
    if (rc) return rc;
    n->launches++;
    mark(n, st, "conv.f " + L.conv_name, 2.0 * L.Cin * L.Cout * L.KS * L.KS * (double)L.Ho * L.Wo * B);
    if (dropping) {
      const Buf& ob = n->bufs[L.out_buf];
      rc = launch_dropout_fwd(wsf(n, ob.act) + L.coff, ob.ld, L.Cout, (int64_t)B * ob.H * ob.W, (int64_t)ob.H * ob.W,
                              n->drop_masks + (size_t)B * L.drop_cprefix, drop_sum, drop_sumsq, st);
      if (rc) return rc;
      n->launches++;
      mark(n, st, "dropout.f " + L.conv_name);
    }
  }
  if (n->coupling) {
    rc = launch_zeros_fwd(out, n->p + n->zb_off, n->p + n->zs_off, B, n->cfg.out_channels, n->out_hw * n->out_hw,
                          tr ? wsf(n, n->out_keep) : nullptr, st);
    if (rc) return rc;
    n->launches++;
    mark(n, st, "conv2dzeros gain");
  }
  if (tr) {
    rc = launch_bn_running_update(bn_table(n), n->n_bn, n->maxC, 0.1f, B, st);
    if (rc) return rc;
    n->launches++;
    mark(n, st, "bn_running_update");
    n->last_B = B;
    n->fwd_train_done = true;
  } else {
    n->fwd_train_done = false;  // the evaluation pass overwrote the activations a backward would need
  }
  return PDES_OK;
}

static int backward_impl(pdes_net_t* n, const float* dout, void* stream, float* dx = nullptr) {
  PDES_REQUIRE(n && n->ws, PDES_ERR_STATE, "pdes_densenet_backward: bind() first");
  PDES_REQUIRE(n->fwd_train_done, PDES_ERR_STATE,
               "pdes_densenet_backward: needs a preceding training-mode forward");
  PDES_REQUIRE(n->g != nullptr, PDES_ERR_STATE, "pdes_densenet_backward: no gradient buffer bound");
  PDES_REQUIRE(dout != nullptr, PDES_ERR_INVALID, "pdes_densenet_backward: null dout");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = n->last_B;
  n->launches = 0;
  int rc;
  bool used_wg = false;
  int wg_rr = 0, side_used = 0;
  bool prejoined = false, tail_side = false;
  mark(n, st, "@backward");
  if (n->coupling) {
    // Conv2dZeros backward of the gain: d(conv) = dout * exp(3 scale), d bias, d scale; the rest of the pass
    // then sees dout * gain as the gradient of the last convolution's (planar) output
    rc = launch_zeros_bwd(dout, wsf(n, n->out_keep), n->p + n->zs_off, B, n->cfg.out_channels, n->out_hw * n->out_hw,
                          wsf(n, n->dyg), n->g + n->zb_off, n->g + n->zs_off, st);
    if (rc) return rc;
    n->launches++;
    mark(n, st, "conv2dzeros bwd");
    dout = wsf(n, n->dyg);
  }
  if (n->conv_impl == 0) {
    // dynamic fp16 scale of the last layer's dY pieces: |dout| maximum (the other layers' slices take
    // the running maximum their gradient buffer collected from the dgrad epilogues)
    rc = launch_absmax(dout, (size_t)B * n->cfg.out_channels * n->out_hw * n->out_hw, gmax_slot(n, -1), st);
    if (rc) return rc;
    n->launches++;
    mark(n, st, "absmax dout");
  }
  for (int li = (int)n->layers.size() - 1; li >= 0; --li) {
    const Layer& L = n->layers[li];
    const bool use_dense_bwd = n->conv_impl == 0 && L.dense_bwd && (n->tc_mask & 2) && (n->tc_mask & 4) &&
                               L.in_buf >= 0 && L.out_buf >= 0;
    const float* dy;
    int lddy = 0, dy_nchw = 0;
    FixDyArgs fixargs;
    bool have_fix = false;
    if (L.out_buf < 0) {
      dy = dout;
      dy_nchw = 1;
    } else {
      const Buf& ob = n->bufs[L.out_buf];
      FixDyArgs f;
      memset(&f, 0, sizeof(f));
      f.G = wsf(n, ob.grad) + L.coff;
      f.X = wsf(n, ob.act) + L.coff;
      f.ldG = f.ldX = ob.ld;
      f.C = L.Cout;
      f.npix = (int64_t)B * ob.H * ob.W;
      f.sum = wsd(n, ob.stat) + L.coff;
      f.sumsq = wsd(n, ob.stat) + ob.C + L.coff;
      f.inv_count = 1.0 / ((double)B * ob.H * ob.W);
      f.eps = 1e-5f;
      for (const auto& M : n->layers) {
        if (M.in_buf != L.out_buf || M.Cin <= L.coff) continue;
        PDES_REQUIRE(f.n_cons < kMaxConsumers, PDES_ERR_UNSUPPORTED, "too many consumers of one tensor");
        f.cons_gamma[f.n_cons] = n->p + M.g_off + L.coff;
        f.cons_bsum[f.n_cons] = wsd(n, M.bsum) + L.coff;
        f.cons_C[f.n_cons] = M.Cin;
        f.n_cons++;
      }
      const bool dropping = n->drop_masks != nullptr && L.drop;
      if (dropping) {
        f.drop_mask = n->drop_masks + (size_t)B * L.drop_cprefix;
        f.pix_per_img = (int64_t)ob.H * ob.W;
      }
      const bool fuse_fix = !dropping && n->conv_impl == 0 && (use_dense_bwd || (L.tc_wg && (n->tc_mask & 4)) ||
                                                  (L.tc2_bwd && !L.dense_bwd && (n->tc_mask & 2)));
      if (!fuse_fix) {
        rc = launch_fix_dy(f, st);
        if (rc) return rc;
        n->launches++;
        mark(n, st, "fix_dy " + L.conv_name);
      } else {
        fixargs = f;
        have_fix = true;
      }
      dy = f.G;
      lddy = ob.ld;
    }
    // ---- wgrad ---------------------------------------------------------------------
    auto run_wgrad = [&]() -> int {
      int rc = PDES_OK;
      WgradArgs w;
      memset(&w, 0, sizeof(w));
      w.B = B;
      w.in_mode = L.up ? IN_UPSAMPLE : IN_DIRECT;
      w.Cin = L.Cin;
      if (L.in_buf >= 0) {
        const Buf& ib = n->bufs[L.in_buf];
        w.x = wsf(n, ib.act);
        w.ldx = ib.ld;
        w.Hs = ib.H;
        w.Ws = ib.W;
        w.pro = 1;
        w.bn = bn_src(n, L, B, true);
      }
      w.dy = dy;
      w.lddy = lddy;
      w.dy_nchw = dy_nchw;
      w.Cout = L.Cout;
      w.KS = L.KS;
      w.pad = L.pad;
      w.stride = L.stride;
      w.Ho = L.Ho;
      w.Wo = L.Wo;
      w.dw = conv_dw(n, L);
      if (L.in_buf < 0) {
        w.x = wsf(n, n->xin);  // NCHW copy made by the training forward
        w.in_nchw = 1;
        w.Hs = L.Hs;
        w.Ws = L.Ws;
      }
      const bool use_wg = n->conv_impl == 0 && L.tc_wg && (n->tc_mask & 4);
      if (L.bilinear && !use_wg) {
        // CUDA-core weight gradient of a bilinear layer: its operand is the upsampled activation, recomputed
        // into the scratch (the tensor-core path re-reads the operand planes the forward pass kept)
        const Buf& ib = n->bufs[L.in_buf];
        BilinearArgs ba;
        memset(&ba, 0, sizeof(ba));
        ba.zero_insert = L.convT ? 1 : 0;
        ba.x = w.x;
        ba.ldx = w.ldx;
        ba.C = L.Cin;
        ba.H = ib.H;
        ba.W = ib.W;
        ba.B = B;
        ba.pro = 1;
        ba.bn = w.bn;
        ba.up = wsf(n, n->up_scratch);
        ba.ldu = rup(L.Cin, 4);
        rc = launch_bilinear_up(ba, st);
        if (rc) return rc;
        n->launches++;
        w.x = ba.up;
        w.ldx = ba.ldu;
        w.Hs = 2 * ib.H;
        w.Ws = 2 * ib.W;
        w.pro = 0;
        w.in_mode = IN_DIRECT;
      }
      // (a layer with the fused dgrad has no conv_tc2 dgrad filter packed: without both backward bits it
      // falls back to the CUDA-core kernels)
      const bool use_dg = n->conv_impl == 0 && L.tc2_bwd && !L.dense_bwd && (n->tc_mask & 2) && L.in_buf >= 0;
      // data gradient as one 1x1 GEMM over the expanded dY planes (needs the weight gradient's expansion)
      const bool dg_i2c = use_dg && use_wg && L.dg_im2col && L.wg_taps_n && !have_fix;
      if ((use_wg || use_dg) && !use_dense_bwd && !(dg_i2c && L.wg_taps_n)) {
        // fp16 pieces of the corrected dY slice: GEMM-K operand of dgrad, GEMM-N operand of wgrad
        ActSplitArgs sb;
        memset(&sb, 0, sizeof(sb));
        sb.x = dy;
        sb.ldx = lddy;
        sb.nchw = dy_nchw;
        sb.C = L.Cout;
        sb.Hs = L.Ho;
        sb.Ws = L.Wo;
        sb.B = B;
        sb.up = L.stride == 2 ? 2 : 0;  // zero-insert: stride-2 layers run as stride-1 kernels
        sb.out = reinterpret_cast<op16*>(wsf(n, L.planesB));
        sb.Cp = (L.Cout + 7) & ~7;
        sb.scale = 1.f;
        sb.dyn_max = gmax_slot(n, L.out_buf);
        sb.dyn_inv = wsf(n, n->dyinv) + li;
        sb.lowp = n->lowp;
        if (have_fix) {
          sb.fix = 1;
          sb.fx = fixargs;
        }
        rc = launch_act_split(sb, st);
        if (rc) return rc;
        n->launches++;
        mark(n, st, "split.b " + L.conv_name);
      }
      if (use_wg && L.wg_taps_n) {
        DyIm2colArgs ia;
        memset(&ia, 0, sizeof(ia));
        ia.dy = dy;
        ia.out = reinterpret_cast<op16*>(wsf(n, L.planesI));
        ia.B = B;
        ia.H = L.Ho;
        ia.W = L.Wo;
        ia.Cout = L.Cout;
        ia.KS = L.KS;
        ia.pad = L.pad;
        ia.Np = (L.KS * L.KS * L.Cout + 7) & ~7;
        ia.dyn_max = gmax_slot(n, L.out_buf);
        ia.dyn_inv = wsf(n, n->dyinv) + li;  // the same value the dY split publishes
        ia.lowp = n->lowp;
        cudaStream_t ist = st;
        if (n->n_side > 0 && n->side[0] && !dg_i2c) {
          // off the critical path: the expansion and the GEMM that reads it both run on the side stream
          // (wg_rr is advanced by the wgrad launch below, which therefore picks the same stream)
          const int k = wg_rr % n->n_side;
          PDES_CUDA(cudaEventRecord(n->ev_fork[k], st));
          PDES_CUDA(cudaStreamWaitEvent(n->side[k], n->ev_fork[k], 0));
          ist = n->side[k];
          side_used |= 1 << k;
        }
        rc = launch_dy_im2col(ia, ist);
        if (rc) return rc;
        n->launches++;
        mark(n, st, "im2col.b " + L.conv_name);
      }
      if (use_wg) {
        const Buf& ib = n->bufs[L.in_buf];
        TcWgradArgs tw;
        memset(&tw, 0, sizeof(tw));
        tw.planesA = reinterpret_cast<const op16*>(wsf(n, L.planes));
        tw.planesB = reinterpret_cast<const op16*>(wsf(n, L.planesB));
        tw.dwp = wsf(n, L.dwp);
        tw.B = B;
        tw.Hv = L.up ? 2 * ib.H : ib.H;
        tw.Wv = L.up ? 2 * ib.W : ib.W;
        tw.Ho = (L.stride == 2 ? 2 : 1) * L.Ho;
        tw.Wo = (L.stride == 2 ? 2 : 1) * L.Wo;
        tw.Cin = L.Cin;
        tw.Cout = L.Cout;
        tw.KS = L.KS;
        tw.pad = L.pad;
        tw.ci_pad = L.ci_pad;
        tw.co_pad = L.co_pad;
        tw.out_scale = pow2f(-kActScaleLog2);
        tw.dyn_scale = wsf(n, n->dyinv) + li;
        tw.lowp = n->lowp;
        if (L.wg_taps_n) {  // one 1x1 GEMM: N = (tap, co) over the expanded dY
          tw.planesB = reinterpret_cast<const op16*>(wsf(n, L.planesI));
          tw.Cout = L.KS * L.KS * L.Cout;
          tw.KS = 1;
          tw.pad = 0;
          tw.co_pad = (tw.Cout + 15) / 16 * 16;
        }
        if (n->n_side > 0 && n->side[0]) {
          // fork: the wgrad kernel only needs the dY planes just written; it runs beside the dgrad chain
          const int k = wg_rr++ % n->n_side;
          PDES_CUDA(cudaEventRecord(n->ev_fork[k], st));
          PDES_CUDA(cudaStreamWaitEvent(n->side[k], n->ev_fork[k], 0));
          rc = launch_wgrad_tc(tw, n->side[k]);
          side_used |= 1 << k;
        } else {
          rc = launch_wgrad_tc(tw, st);
        }
        used_wg = true;
      } else if (n->conv_impl == 0 && L.first_k) {
        FirstConvArgs fa;
        memset(&fa, 0, sizeof(fa));
        fa.x = w.x;
        fa.dy = w.dy;
        fa.lddy = w.lddy;
        fa.dw = w.dw;
        fa.B = B;
        fa.Cin = L.Cin;
        fa.H = L.Hs;
        fa.W = L.Ws;
        fa.Cout = L.Cout;
        fa.KS = L.KS;
        fa.stride = L.stride;
        fa.pad = L.pad;
        fa.Ho = L.Ho;
        fa.Wo = L.Wo;
        if (n->n_side > 0 && n->side[0] && !n->timing) {
          // The last weight gradient of the pass.  Every tensor-core weight gradient issued so far is joined
          // first (pre-join events), so that the unpack of their staging buffers runs on the main stream
          // BESIDE this kernel instead of behind it.
          for (int k = 0; k < 2; ++k)
            if (side_used & (1 << k)) PDES_CUDA(cudaEventRecord(n->ev_join[k], n->side[k]));
          prejoined = true;
          PDES_CUDA(cudaEventRecord(n->ev_fork[0], st));
          PDES_CUDA(cudaStreamWaitEvent(n->side[0], n->ev_fork[0], 0));
          rc = launch_first_conv_wgrad(fa, n->side[0]);
          tail_side = true;
        } else {
          rc = launch_first_conv_wgrad(fa, st);
        }
      } else {
        rc = launch_wgrad_simt(w, st);
      }
      if (rc) return rc;
      n->launches++;
      mark(n, st, "wgrad " + L.conv_name, 2.0 * L.Cin * L.Cout * L.KS * L.KS * (double)L.Ho * L.Wo * B);
      return PDES_OK;
    };
    // ---- dgrad (not needed for the first conv: the input does not require grad) ------
    auto run_dgrad = [&]() -> int {
      int rc = PDES_OK;
      if (L.in_buf < 0) return PDES_OK;
      const Buf& ib = n->bufs[L.in_buf];
      ConvArgs a;
      memset(&a, 0, sizeof(a));
      a.x = dy;
      a.ldx = lddy;
      a.in_nchw = dy_nchw;
      a.Cin = L.Cout;
      a.Hs = L.Ho;
      a.Ws = L.Wo;
      a.B = B;
      a.in_mode = L.stride == 2 ? IN_ZEROINS : IN_DIRECT;
      a.w = wsf(n, L.wb);
      a.CinP = L.CoutPb;
      a.CoP = L.CiPb;
      a.Cout = L.Cin;
      a.KS = L.KS;
      a.pad = L.KS - 1 - L.pad;
      a.stride = 1;
      a.Ho = L.up ? 2 * ib.H : ib.H;
      a.Wo = L.up ? 2 * ib.W : ib.W;
      a.epi = EPI_BNBWD;
      a.pool = L.up;
      a.fx = wsf(n, ib.act);
      if (L.bilinear) {
        // the data gradient w.r.t. the UPSAMPLED activation goes to the scratch as it is; bilinear_bwd gathers
        // it through the transposed interpolation and does the BatchNorm-backward bookkeeping
        a.epi = EPI_NHWC;
        a.pool = 0;
        a.y = wsf(n, n->up_scratch);
        a.ldy = rup(L.Cin, 4);
        a.coff = 0;
      }
      a.ldfx = ib.ld;
      a.Hf = ib.H;
      a.Wf = ib.W;
      a.fbn = bn_src(n, L, B, true);
      a.G = wsf(n, ib.grad);
      a.ldG = ib.ld;
      a.g_accum = L.last_consumer ? 0 : 1;
      a.bsum = wsd(n, L.bsum);
      a.gmax = gmax_slot(n, L.in_buf);
      if (use_dense_bwd) {
        // thin layer: lazy dY correction + dynamic scale + split happen inside the dgrad kernel, which also
        // emits the dY planes of the weight gradient (launched right behind it)
        DenseBwdArgs db;
        memset(&db, 0, sizeof(db));
        db.fx = fixargs;
        db.H = ib.H;
        db.W = ib.W;
        db.B = B;
        db.Cout = L.Cout;
        db.dyn_max = gmax_slot(n, L.out_buf);
        db.dyn_inv = wsf(n, n->dyinv) + li;
        db.planesB = reinterpret_cast<op16*>(wsf(n, L.planesB));
        db.wpk = reinterpret_cast<const op16*>(wsf(n, L.wdb));
        db.N = L.Nb;
        db.Cin = L.Cin;
        db.x = a.fx;
        db.ldx = a.ldfx;
        db.fbn = a.fbn;
        db.G = a.G;
        db.ldG = a.ldG;
        db.g_accum = a.g_accum;
        db.bsum = a.bsum;
        db.gmax = a.gmax;
        db.out_scale = pow2f(-kWScaleLog2);
        db.lowp = n->lowp;
        db.b_early = 1;  // packed during the forward pass
        {
          // PDES_DENSE_DBG_BWD=<layer index>: phase timestamps of that layer's fused dgrad
          static int dbg_layer = -2;
          static long long* dbg_buf = nullptr;
          if (dbg_layer == -2) {
            const char* e = getenv("PDES_DENSE_DBG_BWD");
            dbg_layer = e ? atoi(e) : -1;
            if (dbg_layer >= 0 && cudaMalloc((void**)&dbg_buf, sizeof(long long) * 256) != cudaSuccess) dbg_layer = -1;
          }
          if (dbg_layer == li && dbg_buf != nullptr && !stream_capturing(st)) {
            cudaMemsetAsync(dbg_buf, 0, sizeof(long long) * 256, st);
            db.dbg = dbg_buf;
            rc = launch_conv_dense_bwd(db, st);
            long long h[256];
            cudaMemcpyAsync(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            for (int c = 0; c < 2; ++c) {
              fprintf(stderr, "[dense bwd dbg] %s CTA %d:", L.conv_name.c_str(), c);
              for (int i = 1; i < 64; ++i)
                if (h[c * 64 + i]) fprintf(stderr, " s%d=+%lld", i, h[c * 64 + i] - h[c * 64]);
              fprintf(stderr, "\n");
            }
          } else {
            rc = launch_conv_dense_bwd(db, st);
          }
        }
      } else if (n->conv_impl == 0 && L.dg_im2col && L.wg_taps_n && L.tc_wg && (n->tc_mask & 2) && (n->tc_mask & 4) &&
                 !have_fix) {
        // (the expanded planes were written by dy_im2col_kernel in run_wgrad, on this stream)
        const int Np = (L.KS * L.KS * L.Cout + 7) & ~7;
        Tc2Args t;
        memset(&t, 0, sizeof(t));
        t.c = a;
        t.c.Cin = Np;
        t.c.KS = 1;
        t.c.pad = 0;
        t.c.in_mode = IN_DIRECT;
        t.wpk = reinterpret_cast<const op16*>(wsf(n, L.w2i));
        t.lowp = n->lowp;
        t.out_scale = pow2f(-kWScaleLog2);
        t.dyn_scale = wsf(n, n->dyinv) + li;
        t.N = L.Nb;
        t.KC = L.p2i.KC;
        t.nchunks = L.p2i.nchunks;
        t.ngroups = L.p2i.ngroups;
        t.S = L.p2i.S;
        t.TS = L.p2i.TS;
        t.AST = L.p2i.AST;
        t.NB = L.p2i.NB;
        t.TPB = L.p2i.TPB;
        rc = launch_conv_tc2(t, reinterpret_cast<const op16*>(wsf(n, L.planesI)), L.Ho, L.Wo, Np, st);
      } else if (n->conv_impl == 0 && L.tc2_bwd && !L.dense_bwd && !L.dg_im2col && (n->tc_mask & 2)) {
        Tc2Args t;
        memset(&t, 0, sizeof(t));
        t.c = a;
        t.wpk = reinterpret_cast<const op16*>(wsf(n, L.w2b));
        t.lowp = n->lowp;
        t.out_scale = pow2f(-kWScaleLog2);
        t.dyn_scale = wsf(n, n->dyinv) + li;
        t.N = L.Nb;
        t.KC = L.p2b.KC;
        t.nchunks = L.p2b.nchunks;
        t.ngroups = L.p2b.ngroups;
        t.S = L.p2b.S;
        t.TS = L.p2b.TS;
        t.AST = L.p2b.AST;
        t.NB = L.p2b.NB;
        t.TPB = L.p2b.TPB;
        if (L.stride == 2) t.c.in_mode = IN_DIRECT;  // the zero insertion is in the dY planes
        const int zi = L.stride == 2 ? 2 : 1;
        rc = launch_conv_tc2(t, reinterpret_cast<const op16*>(wsf(n, L.planesB)), zi * L.Ho,
                             zi * L.Wo, L.Cout, st);
      } else {
        rc = launch_conv_simt(a, st);
      }
      if (rc) return rc;
      n->launches++;
      mark(n, st, "dgrad " + L.conv_name, 2.0 * L.Cin * L.Cout * L.KS * L.KS * (double)L.Ho * L.Wo * B);
      if (L.bilinear) {
        BilinearArgs ba;
        memset(&ba, 0, sizeof(ba));
        ba.zero_insert = L.convT ? 1 : 0;
        ba.x = wsf(n, ib.act);
        ba.ldx = ib.ld;
        ba.C = L.Cin;
        ba.H = ib.H;
        ba.W = ib.W;
        ba.B = B;
        ba.bn = bn_src(n, L, B, true);
        ba.up = wsf(n, n->up_scratch);
        ba.ldu = rup(L.Cin, 4);
        ba.G = wsf(n, ib.grad);
        ba.ldG = ib.ld;
        ba.g_accum = L.last_consumer ? 0 : 1;
        ba.bsum = wsd(n, L.bsum);
        ba.gmax = gmax_slot(n, L.in_buf);
        rc = launch_bilinear_bwd(ba, st);
        if (rc) return rc;
        n->launches++;
        mark(n, st, "bilinear.b " + L.conv_name);
      }
      return PDES_OK;
    };
    // the fused dgrad produces the dY planes its layer's weight gradient reads: it goes first
    if (use_dense_bwd) {
      rc = run_dgrad();
      if (rc) return rc;
      rc = run_wgrad();
      if (rc) return rc;
    } else {
      rc = run_wgrad();
      if (rc) return rc;
      rc = run_dgrad();
      if (rc) return rc;
    }
  }
  for (int k = 0; k < 2; ++k) {
    if (!(side_used & (1 << k))) continue;  // join: the unpack reads every staging gradient
    if (!prejoined) PDES_CUDA(cudaEventRecord(n->ev_join[k], n->side[k]));
    PDES_CUDA(cudaStreamWaitEvent(st, n->ev_join[k], 0));
  }
  if (used_wg) {
    rc = launch_wgrad_unpack(wg_table(n), n->n_wg_bound, n->max_wg_cin, n->max_wg_elems, st);
    if (rc) return rc;
    n->launches++;
    mark(n, st, "wgrad_unpack");
  }
  if (tail_side) {   // the first convolution's weight gradient (side stream 0) joins here
    PDES_CUDA(cudaEventRecord(n->ev_join[0], n->side[0]));
    PDES_CUDA(cudaStreamWaitEvent(st, n->ev_join[0], 0));
  }
  for (const auto& L : n->layers)
    if (L.convT) {
      rc = launch_convt_weight_grad(wsf(n, L.gt), n->g + L.w_off, L.Cin, L.Cout, L.KS, st);
      if (rc) return rc;
      n->launches++;
      mark(n, st, "convT_weight_grad " + L.conv_name);
    }
  rc = launch_bn_param_grad(bn_table(n), n->n_bn, n->maxC, st);
  if (rc) return rc;
  n->launches++;
  mark(n, st, "bn_param_grad");
  if (n->coupling && dx != nullptr) {
    // gradient w.r.t. the network input = the first channels of the block's gradient buffer, after the lazy
    // BatchNorm-backward corrections of every layer that normalises those channels
    const Buf& b0 = n->bufs[0];
    FixDyArgs f;
    memset(&f, 0, sizeof(f));
    f.G = wsf(n, b0.grad);
    f.X = wsf(n, b0.act);
    f.ldG = f.ldX = b0.ld;
    f.C = n->cfg.in_channels;
    f.npix = (int64_t)B * b0.H * b0.W;
    f.sum = wsd(n, b0.stat);
    f.sumsq = wsd(n, b0.stat) + b0.C;
    f.inv_count = 1.0 / ((double)B * b0.H * b0.W);
    f.eps = 1e-5f;
    for (const auto& M : n->layers) {
      if (M.in_buf != 0) continue;
      PDES_REQUIRE(f.n_cons < kMaxConsumers, PDES_ERR_UNSUPPORTED, "too many consumers of one tensor");
      f.cons_gamma[f.n_cons] = n->p + M.g_off;
      f.cons_bsum[f.n_cons] = wsd(n, M.bsum);
      f.cons_C[f.n_cons] = M.Cin;
      f.n_cons++;
    }
    rc = launch_fix_dy(f, st);
    if (rc) return rc;
    rc = launch_block_to_nchw(f.G, b0.ld, f.C, B, b0.H * b0.W, dx, st);
    if (rc) return rc;
    n->launches += 2;
    mark(n, st, "input gradient");
  }
  n->fwd_train_done = false;
  return PDES_OK;
}

// ---------------------------------------------------------------------------------------
// public entry points: eager on the first call of a (batch, mode), captured into a CUDA graph on the
// second, replayed afterwards.  When the caller's stream is itself being captured (e.g. the whole
// training step in one outer graph) the launches are issued directly.
// ---------------------------------------------------------------------------------------
namespace {
bool stream_is_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return cs != cudaStreamCaptureStatusNone;
}
// direct launches when: graphs are off, the caller is capturing already (the training engine), per-launch timing
// is on, the network draws dropout masks (a new mask tensor per call) or a kernel debug knob is set
bool exec_graph_ok(const pdes_net* n, cudaStream_t st) {
  static int dbg = -1;
  if (dbg < 0) dbg = (getenv("PDES_DENSE_DBG") || getenv("PDES_DENSE_DBG_BWD")) ? 1 : 0;
  return n->use_graph && !dbg && !n->timing && !n->cfg.dropout && !stream_is_capturing(st);
}
template <class F>
int capture_into(pdes_net::GraphSlot& gs, pdes_net* n, F&& body) {
  cudaStream_t st = n->cap_stream;
  if (st == nullptr || cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  const int rc = body((void*)st);
  cudaGraph_t g = nullptr;
  const cudaError_t e = cudaStreamEndCapture(st, &g);
  if (rc != PDES_OK || e != cudaSuccess || g == nullptr) {
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    return -1;
  }
  cudaGraphExec_t ex = nullptr;
  const cudaError_t ei = cudaGraphInstantiate(&ex, g, 0);
  cudaGraphDestroy(g);
  if (ei != cudaSuccess || ex == nullptr) {
    cudaGetLastError();
    return -1;
  }
  gs.exec = ex;
  gs.launches = n->launches;
  return 0;
}
}  // namespace

extern "C" int pdes_densenet_forward(pdes_net_t* n, const float* x, float* out, int B, int training,
                                     void* stream) {
  PDES_REQUIRE(n && n->ws, PDES_ERR_STATE, "pdes_densenet_forward: bind() first");
  PDES_REQUIRE(x && out, PDES_ERR_INVALID, "pdes_densenet_forward: null pointer");
  PDES_REQUIRE(B >= 1 && B <= n->cfg.max_batch, PDES_ERR_INVALID,
               "pdes_densenet_forward: batch %d outside 1..%d", B, n->cfg.max_batch);
  cudaStream_t st = (cudaStream_t)stream;
  if (!exec_graph_ok(n, st)) return forward_impl(n, x, out, B, training, stream);
  const int tr = training != 0;
  pdes_net::GraphSlot* gs = nullptr;
  for (auto& s : n->gfwd)
    if (s.B == B && s.training == tr) gs = &s;
  if (!gs) {
    gs = &n->gfwd[0];
    for (auto& s : n->gfwd)
      if (s.calls < gs->calls) gs = &s;
    if (gs->exec) cudaGraphExecDestroy(gs->exec);
    *gs = pdes_net::GraphSlot();
    gs->B = B;
    gs->training = tr;
  }
  gs->calls++;
  if (gs->calls == 1) return forward_impl(n, x, out, B, training, stream);  // warm-up (lazy attributes)
  const size_t xbytes = sizeof(float) * (size_t)B * n->cfg.in_channels * n->in_hw * n->in_hw;
  const size_t obytes = sizeof(float) * (size_t)B * n->cfg.out_channels * n->out_hw * n->out_hw;
  if (!gs->exec) {
    pdes_net* nn = n;
    const int rc = capture_into(*gs, n, [&](void* cs) {
      return forward_impl(nn, wsf(nn, nn->xs), wsf(nn, nn->outs), B, training, cs);
    });
    if (rc != 0) {  // capture unavailable: stay on direct launches
      n->use_graph = 0;
      return forward_impl(n, x, out, B, training, stream);
    }
  }
  PDES_CUDA(cudaMemcpyAsync(wsf(n, n->xs), x, xbytes, cudaMemcpyDeviceToDevice, st));
  PDES_CUDA(cudaGraphLaunch(gs->exec, st));
  PDES_CUDA(cudaMemcpyAsync(out, wsf(n, n->outs), obytes, cudaMemcpyDeviceToDevice, st));
  n->launches = gs->launches + 2;
  n->fwd_train_done = tr != 0;
  if (tr) n->last_B = B;
  return PDES_OK;
}

// Backward of a coupling network (arch 2) that also returns the gradient w.r.t. the network input
// (dx: planar (B, in_channels, H, W), overwritten; may be NULL).  Direct launches (no executor graph).
extern "C" int pdes_densenet_backward_dx(pdes_net_t* n, const float* dout, float* dx, void* stream) {
  PDES_REQUIRE(n && n->ws, PDES_ERR_STATE, "pdes_densenet_backward_dx: bind() first");
  PDES_REQUIRE(dx == nullptr || n->coupling, PDES_ERR_UNSUPPORTED,
               "pdes_densenet_backward_dx: the input gradient is implemented for coupling networks (arch 2) only");
  return backward_impl(n, dout, stream, dx);
}

extern "C" int pdes_densenet_backward(pdes_net_t* n, const float* dout, void* stream) {
  PDES_REQUIRE(n && n->ws, PDES_ERR_STATE, "pdes_densenet_backward: bind() first");
  PDES_REQUIRE(n->fwd_train_done, PDES_ERR_STATE,
               "pdes_densenet_backward: needs a preceding training-mode forward");
  PDES_REQUIRE(n->g != nullptr, PDES_ERR_STATE, "pdes_densenet_backward: no gradient buffer bound");
  PDES_REQUIRE(dout != nullptr, PDES_ERR_INVALID, "pdes_densenet_backward: null dout");
  cudaStream_t st = (cudaStream_t)stream;
  if (!exec_graph_ok(n, st)) return backward_impl(n, dout, stream);
  const int B = n->last_B;
  pdes_net::GraphSlot* gs = nullptr;
  for (auto& s : n->gbwd)
    if (s.B == B) gs = &s;
  if (!gs) {
    gs = &n->gbwd[0];
    for (auto& s : n->gbwd)
      if (s.calls < gs->calls) gs = &s;
    if (gs->exec) cudaGraphExecDestroy(gs->exec);
    *gs = pdes_net::GraphSlot();
    gs->B = B;
  }
  gs->calls++;
  if (gs->calls == 1) return backward_impl(n, dout, stream);
  const size_t obytes = sizeof(float) * (size_t)B * n->cfg.out_channels * n->out_hw * n->out_hw;
  if (!gs->exec) {
    pdes_net* nn = n;
    const int rc = capture_into(*gs, n, [&](void* cs) { return backward_impl(nn, wsf(nn, nn->douts), cs); });
    n->fwd_train_done = true;  // the captured body cleared it on the host
    if (rc != 0) {
      n->use_graph = 0;
      return backward_impl(n, dout, stream);
    }
  }
  PDES_CUDA(cudaMemcpyAsync(wsf(n, n->douts), dout, obytes, cudaMemcpyDeviceToDevice, st));
  PDES_CUDA(cudaGraphLaunch(gs->exec, st));
  n->launches = gs->launches + 1;
  n->fwd_train_done = false;
  return PDES_OK;
}
