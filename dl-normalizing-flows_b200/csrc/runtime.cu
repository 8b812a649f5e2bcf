// Host-side runtime of the RealNVP hot path: the plan (topology + parameter table), the
// workspace layout, and the launch schedules for one coupling and for the whole multi-scale
// stack, forward / inverse / backward.  Everything is issued on the caller's stream; nothing
// here synchronises, allocates per call or touches the host after rnvp_plan_bind, so a whole
// step can be captured in a CUDA graph by the caller.
#include <string>
#include <vector>
#include <algorithm>
#include "kernels.h"

#include <mutex>
#include <map>
namespace rnvp {
const char* get_error();
unsigned long long launch_count();

// ---- per-kernel-class event timing --------------------------------------------------
struct ProfRec { int kind, S, taps, cin, cout; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_ev_pool;
static std::mutex g_prof_mu;
static cudaEvent_t get_event() {
  if (!g_ev_pool.empty()) { cudaEvent_t e = g_ev_pool.back(); g_ev_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
ProfScope::ProfScope(int kind, int S, int taps, int cin, int cout, cudaStream_t st_) : slot(-1), st(st_) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{kind, S, taps, cin, cout, get_event(), get_event()};
  cudaEventRecord(r.a, st);
  slot = (int)g_prof.size();
  g_prof.push_back(r);
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[slot].b, st);
}
}
using namespace rnvp;

// =====================================================================================
// plan
// =====================================================================================
namespace {

struct ConvDesc {
  int cin, cout, taps;
  bool has_bias, train_g;
  bool is_skip;          // in_skip / core_skips: a column block of the coupling's fused skip matrix
  int kpad, npad;        // forward operand  wf [taps][npad][kpad]
  int kpad_b, npad_b;    // dgrad operand    wb [taps][npad_b][kpad_b]
  size_t wf_off, wb_off; // float offsets inside the weight arena
  size_t dw_off;         // float offset inside the per-coupling wgrad scratch
  int ld_f, ld_dw;       // row strides of wf / dw in floats (skip convs: column blocks of the fused skip matrix)
  int slot_v, slot_g, slot_bias;   // slots relative to the coupling
};
struct BnDesc {
  int C;
  int slot_w, slot_b, slot_rm, slot_rv;
  size_t sf, sb;         // double offsets of forward / backward sums (2C each)
  size_t save;           // float offset of (mean, rstd, scale, shift) [4C]
};
struct CouplingDesc {
  std::string name;
  int kind, C, S, D, cfg;          // kind 0 = checkerboard, 1 = channelwise
  int cio, cin, cin_pad, cst, cst_pad, ldD;
  std::vector<ConvDesc> convs;     // in_block, in_skip, [rb0, rb3, rb6, core_skip] x R, out
  std::vector<BnDesc> bns;         // [bn1, bn2, bn3] x R, out_block.0
  size_t sf_in, sf_out, sb_cpl, sb_in;   // double offsets
  size_t save_in, save_out;              // float offsets
  int job0;                              // first weight-norm job
  size_t dw_floats;                      // wgrad scratch this coupling needs
  // fused skip path: skip = sum_i skip_conv_i(a_i) is ONE 1x1 conv over the K-concatenated trunk tensors a_0..a_R
  // with the weight matrix wskip [npad][(R+1)*ldD] (every skip conv's wf is a column block of it)
  size_t wskip_off, bskip_off;           // weight arena: the matrix, the summed bias [ldD]
  size_t dwskip_off, dbskip_off;         // wgrad scratch: its gradient, the (shared) bias gradient [ldD]
  CplGeom geom(int B) const {
    CplGeom g;
    g.B = B; g.S = S; g.C = C; g.cio = cio;
    g.ckbd = kind == 0; g.cfg = cfg;
    if (kind == 0) { g.on_off = 0; g.in_off = 0; }
    else if (cfg) { g.on_off = 0; g.in_off = C / 2; }      // modules_realnvp.py:333-336
    else { g.on_off = C / 2; g.in_off = 0; }
    g.cin_pad = cin_pad; g.cst_pad = cst_pad;
    return g;
  }
};

enum { SLOT_SCALE = 0, SLOT_SSHIFT, SLOT_INBN_W, SLOT_INBN_B, SLOT_INBN_RM, SLOT_INBN_RV,
       SLOT_OUTBN_RM, SLOT_OUTBN_RV, SLOT_CONV0 };

constexpr int kMaxR = 16;

// float offsets (relative to the activation base of one coupling) of everything a coupling keeps
struct CplAct {
  size_t h0, a[kMaxR + 1], u1[kMaxR], u2[kMaxR], skip, st, xprime, y, total;
  size_t h[3 * kMaxR + 1];     // mode 2: normalised activations kept for the backward (one per BN)
  bool keep_h;
};

struct Layout {
  size_t weights, dw, accum, stats_f, stats_b, saves, flow, act, scratch, total;   // byte offsets
  size_t stats_f_bytes, stats_b_bytes;
  std::vector<size_t> cpl_act;       // byte offset of each coupling's activation block
  // flow-level buffers (float counts)
  size_t img;                        // B*C*H*W
  // backward scratch (float counts): `sets` sets of `nbuf` trunk-sized buffers followed by the aux block
  size_t maxPD, maxaux, set_floats;
  int nbuf, sets;
};

}  // namespace

struct rnvp_plan {
  rnvp_config cfg;
  std::vector<CouplingDesc> cpl;
  std::vector<std::string> slot_names;
  int slots_per_coupling = 0;
  std::vector<void*> params, grads;
  size_t weight_floats = 0, dw_floats = 0, stats_f_doubles = 0, stats_b_doubles = 0, save_floats = 0;
  WnJob* d_jobs = nullptr;
  std::vector<WnJob> h_jobs;
  int max_cout = 0;
  Seg* d_segs = nullptr;
  int nsegs = 0;
  BiasSumJob* d_biasjobs = nullptr;  // one per coupling: the summed bias of the fused skip conv
  BnEvalJob* d_bnjobs = nullptr;     // every batch norm of the s/t nets (eval-mode coefficient table)
  int n_bnjobs = 0, max_bn_c = 0;
  int math = RNVP_MATH_FP32;
  bool bound = false;
  bool single = false;               // one stand-alone coupling (rnvp_plan_create_single)
  // data parallel
  DpState dp;
  int& world = dp.world;
  // wgrad side stream (mode 2): weight gradients are off the critical path of the backward pass, so they
  // run on a second, lower-priority stream and fill the SMs / HBM time the dgrad chain leaves idle
  cudaStream_t side = nullptr;
  cudaEvent_t ev_ring[32] = {};
  unsigned ev_next = 0;
  cudaEvent_t ev_done[2] = {nullptr, nullptr};   // side stream finished the coupling that used scratch set 0 / 1
  bool done_valid[2] = {false, false};
  cudaEvent_t ev_join = nullptr;
  // forward bookkeeping for backward
  int saved_batch = -1;
  int saved_mode = 1;
  int saved_coupling = -1;           // >=0: a stand-alone coupling forward was saved
  unsigned long long fwd_gen = 0;    // bumped by every forward / inverse call that (over)writes the workspace state
  std::vector<const float*> x_in;    // input of each coupling in the last training forward
};

namespace {

template <typename T> inline T* P_(const rnvp_plan* p, const CouplingDesc& c, int ci, int slot) {
  (void)c;
  return reinterpret_cast<T*>(p->params[(size_t)ci * p->slots_per_coupling + slot]);
}
inline float* G_(const rnvp_plan* p, int ci, int slot) {
  return reinterpret_cast<float*>(p->grads[(size_t)ci * p->slots_per_coupling + slot]);
}
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

ConvDesc make_conv(int cin, int cout, int k, bool bias, bool train_g, int& slot, std::vector<std::string>& names,
                   const std::string& prefix, bool record) {
  ConvDesc c{};
  c.cin = cin; c.cout = cout; c.taps = k * k; c.has_bias = bias; c.train_g = train_g;
  c.kpad = pad_to(cin, 32); c.npad = pad_to(cout, 16);
  c.kpad_b = pad_to(cout, 32); c.npad_b = pad_to(cin, 16);
  c.slot_v = slot++; c.slot_g = slot++; c.slot_bias = slot++;
  if (record) {
    names.push_back(prefix + ".conv.weight_v");
    names.push_back(prefix + ".conv.weight_g");
    names.push_back(bias ? prefix + ".conv.bias" : std::string());
  }
  return c;
}
BnDesc make_bn(int C, int& slot, std::vector<std::string>& names, const std::string& prefix, bool record) {
  BnDesc b{};
  b.C = C;
  b.slot_w = slot++; b.slot_b = slot++; b.slot_rm = slot++; b.slot_rv = slot++;
  if (record) {
    names.push_back(prefix + ".weight");
    names.push_back(prefix + ".bias");
    names.push_back(prefix + ".running_mean");
    names.push_back(prefix + ".running_var");
  }
  return b;
}

struct SingleSpec { int kind, C, S, D, cfg; };

int build_plan(rnvp_plan* p, const SingleSpec* single = nullptr) {
  const rnvp_config& c = p->cfg;
  RNVP_REQUIRE(c.num_scales >= 1 && c.res_blocks >= 1 && c.res_blocks <= kMaxR,
               "num_scales=%d res_blocks=%d unsupported (res_blocks must be in [1,%d])", c.num_scales,
               c.res_blocks, kMaxR);
  RNVP_REQUIRE(c.base_dim > 0 && c.base_dim % 4 == 0, "base_dim=%d must be a positive multiple of 4", c.base_dim);
  RNVP_REQUIRE(c.channels > 0 && c.image_size > 0, "bad channels/image_size");
  RNVP_REQUIRE(single || c.image_size % (1 << (c.num_scales - 1)) == 0, "image_size=%d not divisible by 2^%d",
               c.image_size, c.num_scales - 1);
  RNVP_REQUIRE(c.prior_scale > 0.f, "prior scale must be positive");
  const int R = c.res_blocks;
  int chan = c.channels, size = c.image_size, dim = c.base_dim;
  auto add = [&](const std::string& name, int kind, int C, int S, int D, int cfg) {
    CouplingDesc d;
    d.name = name; d.kind = kind; d.C = C; d.S = S; d.D = D; d.cfg = cfg;
    d.cio = kind == 0 ? C : C / 2;
    d.cin = kind == 0 ? 2 * C + 1 : C;              // modules_realnvp.py:260, 320
    d.cin_pad = pad_to(d.cin, 32);
    d.cst = 2 * d.cio;
    d.cst_pad = pad_to(d.cst, 32);
    d.ldD = pad_to(D, 32);
    bool rec = p->cpl.empty();
    int slot = SLOT_CONV0;
    if (rec) {
      p->slot_names = {"scale", "scale_shift", "in_bn.weight", "in_bn.bias", "in_bn.running_mean",
                       "in_bn.running_var", "out_bn.running_mean", "out_bn.running_var"};
    }
    const std::string b1 = "block.1";
    d.convs.push_back(make_conv(d.cin, D, 3, true, false, slot, p->slot_names, b1 + ".in_block", rec));
    d.convs.push_back(make_conv(D, D, 1, true, true, slot, p->slot_names, b1 + ".in_skip", rec));
    for (int i = 0; i < R; ++i) {
      std::string cb = b1 + ".core_block." + std::to_string(i);
      d.convs.push_back(make_conv(D, D, 1, false, false, slot, p->slot_names, cb + ".res_block.0", rec));
      d.convs.push_back(make_conv(D, D, 3, false, false, slot, p->slot_names, cb + ".res_block.3", rec));
      d.convs.push_back(make_conv(D, D, 1, true, true, slot, p->slot_names, cb + ".res_block.6", rec));
      d.convs.push_back(make_conv(D, D, 1, true, true, slot, p->slot_names,
                                  b1 + ".core_skips." + std::to_string(i), rec));
    }
    d.convs.push_back(make_conv(D, d.cst, 1, true, true, slot, p->slot_names, b1 + ".out_block.2", rec));
    for (int i = 0; i < R; ++i) {
      std::string cb = b1 + ".core_block." + std::to_string(i);
      d.bns.push_back(make_bn(D, slot, p->slot_names, cb + ".in_block.0", rec));
      d.bns.push_back(make_bn(D, slot, p->slot_names, cb + ".res_block.1", rec));
      d.bns.push_back(make_bn(D, slot, p->slot_names, cb + ".res_block.4", rec));
    }
    d.bns.push_back(make_bn(D, slot, p->slot_names, b1 + ".out_block.0", rec));
    if (rec) p->slots_per_coupling = slot;
    p->cpl.push_back(d);
  };
  if (single) {
    RNVP_REQUIRE(single->kind == 0 || single->C % 2 == 0, "channelwise coupling needs an even channel count");
    add(single->kind == 0 ? "ckbd" : "chan", single->kind, single->C, single->S, single->D, single->cfg ? 1 : 0);
  }
  for (int s = 1; !single && s < c.num_scales; ++s) {
    const int ck[3] = {1, 0, 1}, ch[3] = {0, 1, 0};               // flow_realnvp.py:107-116
    for (int i = 0; i < 3; ++i)
      add("s" + std::to_string(s) + "_ckbd." + std::to_string(i), 0, chan, size, dim, ck[i]);
    for (int i = 0; i < 3; ++i)
      add("s" + std::to_string(s) + "_chan." + std::to_string(i), 1, chan * 4, size / 2, dim * 2, ch[i]);
    chan *= 2; size /= 2; dim *= 2;
  }
  const int fin[4] = {1, 0, 1, 0};                                 // flow_realnvp.py:100-105
  for (int i = 0; !single && i < 4; ++i)
    add("s" + std::to_string(c.num_scales) + "_ckbd." + std::to_string(i), 0, chan, size, dim, fin[i]);

  // arenas independent of the batch size
  size_t wo = 0, sf = 0, sb = 0, sv = 0;
  int job = 0;
  for (auto& d : p->cpl) {
    RNVP_REQUIRE(d.cio <= 256, "coupling %s: %d channels unsupported", d.name.c_str(), d.cio);
    RNVP_REQUIRE(d.ldD / 4 <= 256, "coupling %s: D=%d unsupported (max 1024)", d.name.c_str(), d.D);
    d.job0 = job;
    size_t dwo = 0;
    const int nskip = R + 1, ldskip = nskip * d.ldD;
    const size_t skip_floats = align_up((size_t)d.convs[1].npad * ldskip, 64);
    d.wskip_off = wo; wo += skip_floats;
    d.bskip_off = wo; wo += align_up((size_t)d.ldD, 64);
    d.dwskip_off = dwo; dwo += skip_floats;
    d.dbskip_off = dwo; dwo += align_up((size_t)d.ldD, 64);
    int si = 0;
    for (size_t k = 0; k < d.convs.size(); ++k) {
      ConvDesc& cv = d.convs[k];
      const bool is_skip = k == 1 || (k >= 2 && k < d.convs.size() - 1 && (k - 2) % 4 == 3);
      cv.is_skip = is_skip;
      if (is_skip) {                       // column block si of the fused skip matrix
        cv.wf_off = d.wskip_off + (size_t)si * d.ldD;
        cv.dw_off = d.dwskip_off + (size_t)si * d.ldD;
        cv.ld_f = cv.ld_dw = ldskip;
        ++si;
      } else {
        cv.wf_off = wo; wo += align_up((size_t)cv.taps * cv.npad * cv.kpad, 64);
        cv.dw_off = dwo; dwo += align_up((size_t)cv.taps * cv.npad * cv.kpad, 64);
        cv.ld_f = cv.ld_dw = cv.kpad;
      }
      cv.wb_off = wo; wo += align_up((size_t)cv.taps * cv.npad_b * cv.kpad_b, 64);
      p->max_cout = std::max(p->max_cout, cv.cout);
      ++job;
    }
    d.dw_floats = dwo;
    p->dw_floats = std::max(p->dw_floats, dwo);
    d.sf_in = sf; sf += 2 * d.cio;
    d.sf_out = sf; sf += 2 * d.cio;
    d.sb_cpl = sb; sb += 2 * d.cio + 2;
    d.sb_in = sb; sb += 2 * d.cio;
    d.save_in = sv; sv += align_up(4 * d.cio, 4);
    d.save_out = sv; sv += align_up(2 * d.cio, 4);        // keep every block 16-byte aligned (float4 reads)
    for (auto& b : d.bns) {
      b.sf = sf; sf += 2 * b.C;
      b.sb = sb; sb += 2 * b.C;
      b.save = sv; sv += 4 * b.C;
    }
  }
  p->weight_floats = wo;
  p->stats_f_doubles = sf;
  p->stats_b_doubles = sb;
  p->save_floats = sv;
  size_t total = p->cpl.size() * (size_t)p->slots_per_coupling;
  p->params.assign(total, nullptr);
  p->grads.assign(total, nullptr);
  p->x_in.assign(p->cpl.size(), nullptr);
  return RNVP_OK;
}

// math tiers: the tensor-core kernels serve TF32 (operands rounded to TF32 by their producers) and TF32X3 (3xTF32
// split operands, nothing rounded: the fp32-accurate tensor-core tier); FP32 is the CUDA-core reference tier
inline bool tc_tier(const rnvp_plan* p) { return p->math != RNVP_MATH_FP32; }
inline bool x3_tier(const rnvp_plan* p) { return p->math == RNVP_MATH_TF32X3; }
bool cpl_xf(const rnvp_plan* p, const CouplingDesc& d);
bool bn_fused(const rnvp_plan* p, const CouplingDesc& d, int bi);

CplAct cpl_act(const rnvp_plan* p, const CouplingDesc& d, int B, int mode) {
  CplAct a{};
  const int R = p->cfg.res_blocks;
  size_t Pn = (size_t)B * d.S * d.S, o = 0;
  auto take = [&](size_t n) { size_t r = o; o += align_up(n, 64); return r; };
  a.h0 = take(Pn * d.cin_pad);
  if (mode >= 1) {
    for (int i = 0; i <= R; ++i) a.a[i] = take(Pn * d.ldD);
    for (int i = 0; i < R; ++i) { a.u1[i] = take(Pn * d.ldD); a.u2[i] = take(Pn * d.ldD); }
  } else {                               // inference: u1 / u2 are one scratch buffer; the trunk tensors a_0..a_R stay
    for (int i = 0; i <= R; ++i) a.a[i] = take(Pn * d.ldD);   // (the fused skip conv reads all of them at the end)
    size_t uu = take(Pn * d.ldD);
    for (int i = 0; i < R; ++i) { a.u1[i] = uu; a.u2[i] = uu; }
  }
  a.skip = take(Pn * d.ldD);
  // mode 2 keeps relu(bn(.)) for the backward -- except where the BN is applied inside its consumer kernels
  a.keep_h = mode == 2;
  if (a.keep_h)
    for (int i = 0; i < 3 * R + 1; ++i)
      if (!bn_fused(p, d, i)) a.h[i] = take(Pn * d.ldD);
  a.st = take(Pn * d.cst_pad);
  a.xprime = take(Pn * d.cio);
  a.y = take(Pn * d.C);
  a.total = o;
  return a;
}

Layout compute_layout(const rnvp_plan* p, int B, int mode) {
  Layout L{};
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 1024); return r; };
  L.weights = take(p->weight_floats * 4 * (x3_tier(p) ? 2 : 1));     // 3xTF32: the lo copy follows the weights
  L.dw = take(mode >= 1 ? p->dw_floats * 4 : 0);
  L.accum = take((32 + 2 * (size_t)B) * 8);
  L.stats_f_bytes = p->stats_f_doubles * 8;
  L.stats_b_bytes = p->stats_b_doubles * 8;
  L.stats_f = take(L.stats_f_bytes);
  L.stats_b = take(mode >= 1 ? L.stats_b_bytes : 0);
  L.saves = take(p->save_floats * 4);
  const rnvp_config& c = p->cfg;
  L.img = (size_t)B * c.channels * c.image_size * c.image_size;
  // flow-level buffers (see flow_bufs): persistent group inputs + factored-out halves (< 5 img),
  // two forward temporaries, three gradient temporaries, one per-sample vector
  L.flow = take((L.img * 12 + 64 * p->cfg.num_scales + (size_t)B + 1024) * 4);
  size_t maxact = 0, maxPD = 0, maxaux = 0;
  L.cpl_act.resize(p->cpl.size());
  size_t act_total = 0;
  for (size_t i = 0; i < p->cpl.size(); ++i) {
    const CouplingDesc& d = p->cpl[i];
    CplAct a = cpl_act(p, d, B, mode);
    L.cpl_act[i] = act_total * 4;
    if (mode >= 1) act_total += align_up(a.total, 256);
    maxact = std::max(maxact, a.total);
    size_t Pn = (size_t)B * d.S * d.S;
    maxPD = std::max(maxPD, align_up(Pn * d.ldD, 64));
    maxaux = std::max(maxaux, align_up(Pn * d.cst_pad, 64) + align_up(Pn * d.cio, 64) + align_up(Pn * d.cin_pad, 64));
  }
  if (mode < 1) act_total = align_up(maxact, 256);
  L.act = take(act_total * 4);
  for (auto& v : L.cpl_act) v += L.act;
  // scratch.  inference: H.  mode 1: H, T0, T1, DA, DO + aux (dst/dxdir/dh0), buffers reused layer to layer.
  // mode 2: every gradient tensor of a coupling's s/t net gets its own buffer (5 per residual block + 3), in
  // two sets used by alternate couplings, so that the side-stream wgrads never race with the dgrad chain.
  L.maxPD = maxPD; L.maxaux = maxaux;
  L.nbuf = mode == 2 ? 5 * p->cfg.res_blocks + 3 : (mode == 1 ? 5 : 1);
  L.sets = mode == 2 ? 2 : 1;
  L.set_floats = (size_t)L.nbuf * maxPD + (mode >= 1 ? maxaux : 0);
  L.scratch = take(L.sets * L.set_floats * 4);
  L.total = o;
  return L;
}

// ------------------------------------------------------------------------------------
// execution context of one call
// ------------------------------------------------------------------------------------
struct Ctx {
  rnvp_plan* p;
  Layout L;
  char* ws;
  int B, mode;
  cudaStream_t st;
  cudaStream_t wst;                  // stream of the wgrad kernels (== st unless the side stream is active)
  bool side_on;
  float* weights() const { return reinterpret_cast<float*>(ws + L.weights); }
  float* dw() const { return reinterpret_cast<float*>(ws + L.dw); }
  double* ws_acc() const { return reinterpret_cast<double*>(ws + L.accum); }
  double* logdet_acc() const { return reinterpret_cast<double*>(ws + L.accum) + 32; }
  double* prior_acc() const { return logdet_acc() + B; }
  double* sf(size_t off) const { return reinterpret_cast<double*>(ws + L.stats_f) + off; }
  double* sb(size_t off) const { return reinterpret_cast<double*>(ws + L.stats_b) + off; }
  float* save(size_t off) const { return reinterpret_cast<float*>(ws + L.saves) + off; }
  float* act(int ci, size_t off) const { return reinterpret_cast<float*>(ws + L.cpl_act[ci]) + off; }
  float* scratch() const { return reinterpret_cast<float*>(ws + L.scratch); }
  // backward scratch of coupling ci: trunk-sized buffer k, and the aux block behind the buffers
  float* set_base(int ci) const { return scratch() + (size_t)(L.sets == 2 ? (ci & 1) : 0) * L.set_floats; }
  float* buf(int ci, int k) const { return set_base(ci) + (size_t)k * L.maxPD; }
  float* aux(int ci) const { return set_base(ci) + (size_t)L.nbuf * L.maxPD; }
};

bool side_stream_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("RNVP_WGRAD_STREAM");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

// BatchNorm2d + ReLU folded into the consuming tensor-core kernels (conv_tc.cu "BN prologue"): on by default,
// RNVP_XFORM=0 restores the separate bn_relu passes (A/B measurements)
bool xform_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("RNVP_XFORM");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}
// true when the tensor-core kernels of coupling `d` can apply a BN to their operand tiles (shape / tier check)
bool cpl_xf(const rnvp_plan* p, const CouplingDesc& d) {
  if (!tc_tier(p) || !xform_enabled()) return false;
  ConvArgs a{};
  a.S = d.S; a.kpad = d.ldD; a.ldy = d.ldD; a.n = d.D;
  WgradArgs w{};
  w.S = d.S; w.kpad = d.ldD; w.lddy = d.ldD;
  return conv_tf32_prologue_ok(a) && wgrad_tf32_prologue_ok(w);
}

// BN `bi` of the s/t net ([bn1, bn2, bn3] x R, out_block.0) is folded into its consumer conv and that conv's wgrad,
// i.e. relu(bn(.)) never exists in HBM.  Only the 1x1 consumers (rb0 after bn1, rb6 after bn3, the out conv) take
// the fold: a 3x3 conv would re-transform every activation tile for each of its nine taps -- measured 2x slower than
// one bn_relu pass + the plain kernel (profiles/r02a_*), its operand feed being shared-memory-bandwidth bound.
bool bn_fused(const rnvp_plan* p, const CouplingDesc& d, int bi) {
  const int R = p->cfg.res_blocks;
  const bool consumer_1x1 = bi == 3 * R || bi % 3 != 1;
  return consumer_1x1 && cpl_xf(p, d);
}

// The skip path as ONE conv: skip = sum_i skip_conv_i(a_i) = [a_0 | ... | a_R] x wskip, a 1x1 conv whose K dimension
// runs over the R+1 trunk tensors (a fifth TMA dimension), instead of R+1 launches that each read and re-write the
// running sum: 6 instead of 14 trunk-sized HBM passes per coupling, 4 launches fewer.  Its wgrad is likewise one
// launch over the concatenated x.  Tensor-core tier only; RNVP_SKIP_FUSED=0 restores the separate convs (A/B).
bool skip_fused(const rnvp_plan* p, const CouplingDesc& d, int B, int mode) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("RNVP_SKIP_FUSED");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  if (!on || !tc_tier(p)) return false;
  ConvArgs a{};
  a.S = d.S; a.kpad = d.ldD; a.ldy = d.ldD; a.n = d.D;
  WgradArgs w{};
  w.S = d.S; w.kpad = d.ldD; w.lddy = d.ldD;
  CplAct A = cpl_act(p, d, B, mode);
  return conv_tf32_fusable(a) && wgrad_tf32_prologue_ok(w) && (A.a[1] - A.a[0]) % 4 == 0;
}

// the affine coupling map computed in the epilogue of the s/t net's out conv (needs the BN-prologue kernel);
// RNVP_CPL_EPILOGUE=0 restores the separate cpl_fwd_a / cpl_inv passes (A/B measurements)
bool cpl_fused(const rnvp_plan* p, const CouplingDesc& d) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("RNVP_CPL_EPILOGUE");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0 && cpl_xf(p, d) && d.cst <= 128;
}

int make_ctx(rnvp_plan* p, int B, int mode, void* ws, size_t ws_bytes, void* stream, Ctx* c) {
  RNVP_REQUIRE(p && p->bound, "plan is not bound to parameters (call rnvp_plan_bind)");
  RNVP_REQUIRE(B > 0, "batch must be positive");
  if (const int e = dp_sticky_error(&p->dp)) {
    set_error("data-parallel statistic exchange failed earlier (%s); the replicas have diverged -- rebuild the "
              "communicator and restore a checkpoint",
              e == 1 ? "a peer rank never arrived" : (e == 2 ? "the ranks issued different call sequences"
                                                              : "the ranks' local batch sizes differ"));
    return RNVP_ERR_STATE;
  }
  c->p = p; c->B = B; c->mode = mode; c->st = (cudaStream_t)stream;
  c->side_on = mode == 2 && p->side != nullptr && side_stream_enabled() && !g_prof_on;
  c->wst = c->side_on ? p->side : c->st;
  c->L = compute_layout(p, B, mode);
  if (ws == nullptr || ws_bytes < c->L.total) {
    set_error("workspace too small: need %zu bytes, got %zu", c->L.total, ws_bytes);
    return RNVP_ERR_WORKSPACE;
  }
  RNVP_REQUIRE(((uintptr_t)ws & 255) == 0, "workspace must be 256-byte aligned");
  c->ws = reinterpret_cast<char*>(ws);
  return RNVP_OK;
}

int sync_stats(const Ctx& c, double* buf, size_t n) {
  if (c.p->world > 1) return dp_allreduce_doubles(&c.p->dp, buf, n, c.st);
  return RNVP_OK;
}
// the same reduction folded into the kernel that consumes `buf` (bn_relu, bn_bwd_apply, the BN-prologue conv)
int sync_stats_fused(const Ctx& c, double* buf, size_t n, DpXchg* xg) {
  *xg = DpXchg();
  if (c.p->world > 1) return dp_fused_exchange(&c.p->dp, buf, n, c.st, xg);
  return RNVP_OK;
}

ConvArgs conv_args(const Ctx& c, const ConvDesc& cv, bool dgrad, const float* x, int S, float* y, int ldy,
                   const float* bias, const float* res, double* stats) {
  ConvArgs a{};
  a.x = x;
  a.w = c.weights() + (dgrad ? cv.wb_off : cv.wf_off);
  a.bias = bias; a.res = res; a.y = y; a.stats = stats;
  a.B = c.B; a.S = S;
  a.kpad = dgrad ? cv.kpad_b : cv.kpad;
  a.n = dgrad ? cv.cin : cv.cout;
  a.npad = dgrad ? cv.npad_b : cv.npad;
  a.taps = cv.taps; a.ldy = ldy;
  a.ldw = dgrad ? cv.kpad_b : cv.ld_f;
  if (x3_tier(c.p)) { a.x3 = 1; a.w_lo_delta = c.p->weight_floats; }
  return a;
}

int run_conv(const Ctx& c, const ConvDesc& cv, bool dgrad, const float* x, int S, float* y, int ldy,
             const float* bias, const float* res, double* stats, bool operand_out = false,
             const float* post_save = nullptr /* (mean, rstd, scale, shift)[n] of the next layer's BN: eval fold */) {
  ProfScope ps(dgrad ? PROF_DGRAD : PROF_CONV, S, cv.taps, dgrad ? cv.cout : cv.cin, dgrad ? cv.cin : cv.cout, c.st);
  ConvArgs a = conv_args(c, cv, dgrad, x, S, y, ldy, bias, res, stats);
  // y is read raw by later conv MMAs: round it where it is produced (tensor-core tier only)
  a.round_out = operand_out && c.p->math == RNVP_MATH_TF32;
  if (post_save) { a.post_scale = post_save + 2 * a.n; a.post_shift = post_save + 3 * a.n; }
  return tc_tier(c.p) ? k_conv_fwd_tf32(a, c.st) : k_conv_fwd_fp32(a, c.st);
}
// conv whose input is relu(bn_bi(x_raw)): BN prologue inside the tensor-core kernel
int run_conv_bn(const Ctx& c, int ci, const ConvDesc& cv, int bi, int training, double count, const float* x_raw, int S,
                float* y, int ldy, const float* bias, const float* res, double* stats, bool operand_out,
                const CplEpilogue* cpl = nullptr, const float* post_save = nullptr) {
  rnvp_plan* p = c.p;
  const CouplingDesc& d = p->cpl[ci];
  const BnDesc& b = d.bns[bi];
  ProfScope ps(PROF_CONV, S, cv.taps, cv.cin, cv.cout, c.st);
  ConvArgs a = conv_args(c, cv, false, x_raw, S, y, ldy, bias, res, stats);
  a.round_out = operand_out && p->math == RNVP_MATH_TF32;
  if (post_save) { a.post_scale = post_save + 2 * a.n; a.post_shift = post_save + 3 * a.n; }
  BnPrologue x{};
  x.mode = training ? 1 : 0; x.C = b.C; x.sums = c.sf(b.sf); x.count = count;
  x.gamma = P_<float>(p, d, ci, b.slot_w); x.beta = P_<float>(p, d, ci, b.slot_b);
  x.run_mean = P_<float>(p, d, ci, b.slot_rm); x.run_var = P_<float>(p, d, ci, b.slot_rv);
  x.save = c.save(b.save);
  if (training) RNVP_TRY(sync_stats_fused(c, c.sf(b.sf), 2 * b.C, &x.xg));
  a.xf = &x;
  a.cpl = cpl;
  return k_conv_fwd_tf32(a, c.st);
}
// the side stream waits for everything enqueued on the main stream so far
int fork_to_side(const Ctx& c) {
  if (!c.side_on) return RNVP_OK;
  rnvp_plan* p = c.p;
  cudaEvent_t e = p->ev_ring[p->ev_next++ % 32];
  RNVP_CUDA(cudaEventRecord(e, c.st));
  RNVP_CUDA(cudaStreamWaitEvent(c.wst, e, 0));
  return RNVP_OK;
}
// the main stream waits for everything enqueued on the side stream; forgets the per-set markers
int join_side(const Ctx& c) {
  if (!c.side_on) return RNVP_OK;
  rnvp_plan* p = c.p;
  RNVP_CUDA(cudaEventRecord(p->ev_join, c.wst));
  RNVP_CUDA(cudaStreamWaitEvent(c.st, p->ev_join, 0));
  p->done_valid[0] = p->done_valid[1] = false;
  return RNVP_OK;
}

// xf_save != null: x is the raw pre-BN tensor and the kernel applies relu(bn(.)) to its boxes (tensor-core tier)
int run_wgrad(const Ctx& c, const ConvDesc& cv, const float* x, const float* dy, int lddy, int S, float* dbias,
              const float* xf_save = nullptr, int xf_C = 0) {
  RNVP_TRY(fork_to_side(c));                 // dy (and a recomputed x) were produced on the main stream
  ProfScope ps(PROF_WGRAD, S, cv.taps, cv.cin, cv.cout, c.wst);
  WgradArgs a{};
  a.x = x; a.dy = dy; a.dw = c.dw() + cv.dw_off; a.dbias = dbias;
  a.xf_save = xf_save; a.xf_C = xf_C;
  a.B = c.B; a.S = S; a.kpad = cv.kpad; a.n = cv.cout; a.npad = cv.npad; a.taps = cv.taps; a.lddy = lddy;
  a.lddw = cv.ld_dw;
  a.x3 = x3_tier(c.p);
  return tc_tier(c.p) ? k_conv_wgrad_tf32(a, c.wst) : k_conv_wgrad_fp32(a, c.wst);
}

// ------------------------------------------------------------------------------------
// s/t network (ResidualModule, modules_realnvp.py:175-194) forward
// ------------------------------------------------------------------------------------
// `cpl` (tensor-core tier only): the coupling map fused into the out conv's epilogue
int net_forward(const Ctx& c, int ci, int training, const CplEpilogue* cpl = nullptr) {
  rnvp_plan* p = c.p;
  const CouplingDesc& d = p->cpl[ci];
  const int R = p->cfg.res_blocks, S = d.S, ld = d.ldD, Pn = c.B * S * S;
  const double count = (double)Pn * p->world;
  CplAct A = cpl_act(p, d, c.B, c.mode);
  float* H = c.scratch();
  const int rnd = p->math == RNVP_MATH_TF32;
  auto bias = [&](const ConvDesc& cv) { return cv.has_bias ? P_<float>(p, d, ci, cv.slot_bias) : nullptr; };
  float* Hs = c.scratch();
  auto bn = [&](int bi, const float* x) -> int {
    const BnDesc& b = d.bns[bi];
    DpXchg xg;
    if (training) RNVP_TRY(sync_stats_fused(c, c.sf(b.sf), 2 * b.C, &xg));
    ProfScope ps(PROF_BN, S, 0, b.C, b.C, c.st);
    H = A.keep_h ? c.act(ci, A.h[bi]) : Hs;
    return k_bn_relu(x, H, Pn, b.C, ld, training ? c.sf(b.sf) : nullptr, count, P_<float>(p, d, ci, b.slot_w),
                     P_<float>(p, d, ci, b.slot_b), P_<float>(p, d, ci, b.slot_rm),
                     P_<float>(p, d, ci, b.slot_rv), c.save(b.save), training ? 1 : 0, rnd, c.st, xg);
  };
  auto st_of = [&](int bi) { return training ? c.sf(d.bns[bi].sf) : nullptr; };
  const ConvDesc* cv = d.convs.data();
  const bool xf = bn_fused(p, d, 3 * R);
  // y = conv(relu(bn_bi(x))): one kernel with the BN prologue, or bn_relu + conv
  auto bn_conv = [&](int bi, const float* x, const ConvDesc& cvx, float* y, int ldy, const float* b, const float* res,
                     double* stats, bool operand_out) -> int {
    if (bn_fused(p, d, bi))
      return run_conv_bn(c, ci, cvx, bi, training, count, x, S, y, ldy, b, res, stats, operand_out);
    RNVP_TRY(bn(bi, x));
    return run_conv(c, cvx, false, H, S, y, ldy, b, res, stats, operand_out);
  };
  // The skip path (in_skip and the core_skips, accumulated into `skip`) is off the critical chain of the
  // trunk: with the side stream active (training mode 2, every a_i has its own buffer) those five convs
  // overlap the residual blocks and are joined before out_block's batch norm.
  const bool sfused = skip_fused(p, d, c.B, c.mode);
  auto skip_conv = [&](const ConvDesc& cvs, const float* x, const float* res, double* stats) -> int {
    if (sfused) return RNVP_OK;              // one K-concatenated conv after the last block
    if (!c.side_on) return run_conv(c, cvs, false, x, S, c.act(ci, A.skip), ld, bias(cvs), res, stats);
    RNVP_TRY(fork_to_side(c));
    Ctx cs2 = c;
    cs2.st = c.wst;
    return run_conv(cs2, cvs, false, x, S, c.act(ci, A.skip), ld, bias(cvs), res, stats);
  };
  // a0 = in_block(h0); skip = in_skip(a0)
  RNVP_TRY(run_conv(c, cv[0], false, c.act(ci, A.h0), S, c.act(ci, A.a[0]), ld, bias(cv[0]), nullptr, st_of(0), true));
  RNVP_TRY(skip_conv(cv[1], c.act(ci, A.a[0]), nullptr, nullptr));
  for (int i = 0; i < R; ++i) {
    const ConvDesc *rb0 = &cv[2 + 4 * i], *rb3 = rb0 + 1, *rb6 = rb0 + 2, *cs = rb0 + 3;
    float *ai = c.act(ci, A.a[i]), *an = c.act(ci, A.a[i + 1]);
    if (!training && xf) {
      // eval mode: every BN is a fixed affine map (coefficients in `save`, eval_bn_coefs).  bn2 / bn3 + ReLU ride in the
      // epilogue of the conv that PRODUCES their input (no other consumer needs u1 / u2 raw), bn1 in rb0's operand
      // path: no bn_relu pass at all, and the 3x3 conv reads a ready-made operand.
      float* h2 = Hs;                                   // scratch: the 3x3 conv cannot run in place on u1 == u2
      float* h3 = c.act(ci, A.u2[i]);
      RNVP_TRY(run_conv_bn(c, ci, *rb0, 3 * i, 0, count, ai, S, h2, ld, nullptr, nullptr, nullptr, true, nullptr,
                           c.save(d.bns[3 * i + 1].save)));
      RNVP_TRY(run_conv(c, *rb3, false, h2, S, h3, ld, nullptr, nullptr, nullptr, true, c.save(d.bns[3 * i + 2].save)));
      RNVP_TRY(run_conv(c, *rb6, false, h3, S, an, ld, bias(*rb6), ai, nullptr, true));
      RNVP_TRY(skip_conv(*cs, an, c.act(ci, A.skip), nullptr));
      continue;
    }
    RNVP_TRY(bn_conv(3 * i, ai, *rb0, c.act(ci, A.u1[i]), ld, nullptr, nullptr, st_of(3 * i + 1), false));
    RNVP_TRY(bn_conv(3 * i + 1, c.act(ci, A.u1[i]), *rb3, c.act(ci, A.u2[i]), ld, nullptr, nullptr, st_of(3 * i + 2), false));
    RNVP_TRY(bn_conv(3 * i + 2, c.act(ci, A.u2[i]), *rb6, an, ld, bias(*rb6), ai, i + 1 < R ? st_of(3 * (i + 1)) : nullptr, true));
    RNVP_TRY(skip_conv(*cs, an, c.act(ci, A.skip), i == R - 1 ? st_of(3 * R) : nullptr));
  }
  if (sfused) {
    ProfScope ps(PROF_CONV, S, 1, (R + 1) * d.D, d.D, c.st);
    ConvArgs a = conv_args(c, cv[1], false, c.act(ci, A.a[0]), S, c.act(ci, A.skip), ld, c.weights() + d.bskip_off, nullptr,
                           st_of(3 * R));
    a.w = c.weights() + d.wskip_off;
    a.segs = R + 1; a.seg_stride = A.a[1] - A.a[0]; a.ldw = (R + 1) * ld;
    RNVP_TRY(k_conv_fwd_tf32(a, c.st));
  }
  RNVP_TRY(join_side(c));
  const ConvDesc& oc = cv[2 + 4 * R];
  if (cpl) {
    RNVP_REQUIRE(xf, "internal: coupling epilogue without the BN-prologue kernel");
    return run_conv_bn(c, ci, oc, 3 * R, training, count, c.act(ci, A.skip), S, c.act(ci, A.st), d.cst_pad, bias(oc),
                       nullptr, nullptr, false, cpl);
  }
  RNVP_TRY(bn_conv(3 * R, c.act(ci, A.skip), oc, c.act(ci, A.st), d.cst_pad, bias(oc), nullptr, nullptr, false));
  return RNVP_OK;
}

// backward of the s/t network: dst [P,cst_pad] -> dh0 [P,cin_pad]; weight grads into the dw scratch.
// Gradient buffers come from c.buf(ci, k).  Mode 1 has five of them and reuses them layer to layer
// (H = recomputed activation, T0, T1, DA, DO); mode 2 takes a fresh buffer for every gradient tensor so
// that the wgrad kernels, which run on the side stream, can read their dy operand at any later time.
int net_backward(const Ctx& c, int ci, const float* dst, float* dh0) {
  rnvp_plan* p = c.p;
  const CouplingDesc& d = p->cpl[ci];
  const int R = p->cfg.res_blocks, S = d.S, ld = d.ldD, Pn = c.B * S * S;
  const double count = (double)Pn * p->world;
  CplAct A = cpl_act(p, d, c.B, c.mode);
  const bool fresh = c.mode == 2;
  int next = 0;
  // mode 1 roles: 0 = H, 1 = T0, 2 = T1, 3 = DA, 4 = DO
  auto take = [&](int role) { return fresh ? c.buf(ci, next++) : c.buf(ci, role); };
  float* Hs = fresh ? nullptr : c.buf(ci, 0);
  float* H = Hs;
  auto gbias = [&](const ConvDesc& cv) { return cv.has_bias ? G_(p, ci, cv.slot_bias) : nullptr; };
  // wgrad of a conv whose input was relu(bn_bi(x)): the kernel re-applies the BN to the raw x boxes (folded BNs),
  // else H holds the (kept or recomputed) normalised activation
  auto wgrad_bn = [&](const ConvDesc& cvw, int bi, const float* x, const float* dy, int lddy, float* dbias) -> int {
    if (bn_fused(p, d, bi)) return run_wgrad(c, cvw, x, dy, lddy, S, dbias, c.save(d.bns[bi].save), d.bns[bi].C);
    return run_wgrad(c, cvw, H, dy, lddy, S, dbias);
  };
  auto recompute = [&](int bi, const float* x) -> int {
    const BnDesc& b = d.bns[bi];
    if (bn_fused(p, d, bi)) return RNVP_OK;   // relu(bn(x)) is rebuilt inside the wgrad kernel
    if (A.keep_h) {                      // mode 2: the forward kept relu(bn(x))
      H = c.act(ci, A.h[bi]);
      return RNVP_OK;
    }
    H = Hs;
    ProfScope ps(PROF_BN, S, 0, b.C, b.C, c.st);
    return k_bn_relu(x, H, Pn, b.C, ld, nullptr, count, nullptr, nullptr, nullptr, nullptr, c.save(b.save), 2,
                     p->math == RNVP_MATH_TF32, c.st);
  };
  // dgrad of `cv` applied to `dy`, then ReLU+BN backward through BN `bi` whose raw input was `x`:
  // g <- masked gradient, out = d(pre-BN input) (+ add).  With the tensor-core tier the mask and the two
  // reductions ride in the dgrad epilogue; otherwise a separate reduce kernel does them.
  // `operand_out`: `out` is the dy operand of later dgrad / wgrad MMAs (rounded to TF32 by the apply kernel)
  auto dgrad_bn_bwd = [&](const ConvDesc& cvd, const float* dy, int bi, float* g, const float* x, float* out,
                          const float* add, bool operand_out) -> int {
    const BnDesc& b = d.bns[bi];
    ConvArgs a = conv_args(c, cvd, true, dy, S, g, ld, nullptr, nullptr, nullptr);
    const bool fused = p->math == RNVP_MATH_TF32 && conv_tf32_fusable(a);
    if (fused) {
      ProfScope ps(PROF_DGRAD, S, cvd.taps, cvd.cout, cvd.cin, c.st);
      a.bn_x = x; a.bn_save = c.save(b.save); a.stats = c.sb(b.sb);
      RNVP_TRY(k_conv_fwd_tf32(a, c.st));
    } else {
      RNVP_TRY(run_conv(c, cvd, true, dy, S, g, ld, nullptr, nullptr, nullptr));
    }
    ProfScope ps(PROF_BN_BWD, S, 0, b.C, b.C, c.st);
    if (!fused) RNVP_TRY(k_bn_bwd_reduce(g, x, g, Pn, b.C, ld, c.save(b.save), c.sb(b.sb), c.st));
    DpXchg xg;
    RNVP_TRY(sync_stats_fused(c, c.sb(b.sb), 2 * b.C, &xg));
    return k_bn_bwd_apply(g, x, out, add, Pn, b.C, ld, c.save(b.save), c.sb(b.sb), count,
                          P_<float>(p, d, ci, b.slot_w), G_(p, ci, b.slot_w), G_(p, ci, b.slot_b),
                          1.0f / p->world, fused ? 1 : 0, operand_out && p->math == RNVP_MATH_TF32, c.st, xg);
  };
  const ConvDesc* cv = d.convs.data();
  const ConvDesc& oc = cv[2 + 4 * R];
  // out_block: st = conv(relu(bn(skip)));  DO = d(skip sum), constant for the rest of the coupling
  float* DO = take(4);
  {
    float* G0 = take(1);
    RNVP_TRY(recompute(3 * R, c.act(ci, A.skip)));
    RNVP_TRY(wgrad_bn(oc, 3 * R, c.act(ci, A.skip), dst, d.cst_pad, gbias(oc)));
    RNVP_TRY(dgrad_bn_bwd(oc, dst, 3 * R, G0, c.act(ci, A.skip), DO, nullptr, true));
  }
  // Grouped weight gradients (tensor-core tiers, workspace mode 2 where every dy has its own buffer): the 2R 1x1 wgrads
  // (rb0, rb6 of every block) and the R 3x3 wgrads (rb3) of the coupling are collected here and issued as ONE launch
  // each after the dgrad chain -- 3R - 2 launches fewer per coupling.  RNVP_WGRAD_GROUPED=0 restores one launch per conv.
  // Single process only: under data parallelism the grouped launches measured SLOWER (2 GPUs, same session: 73.1 - 73.4
  // vs 72.2 ms per step) -- one long low-priority kernel holds the SMs while the main stream's statistic exchanges and
  // their small consumers queue behind it, and every such delay is a wait for all ranks at the next exchange.
  static const bool grouped_on = [] { const char* e = getenv("RNVP_WGRAD_GROUPED"); return !(e && e[0] == '0'); }();
  // TF32 tier only: the 3xTF32 wgrad runs one CTA per SM on a shallow ring, and fewer pixel splits per job cost it
  // more than the launches save (measured 136.9 vs 129.2 ms per step).
  bool grouped = grouped_on && p->world == 1 && p->math == RNVP_MATH_TF32 && fresh && A.keep_h && 2 * R <= kMaxWgradJobs;
  for (int i = 0; i < R && grouped; ++i) grouped = bn_fused(p, d, 3 * i) && bn_fused(p, d, 3 * i + 2) && !bn_fused(p, d, 3 * i + 1);
  {
    WgradArgs w{};
    w.S = S; w.kpad = ld; w.lddy = ld;
    grouped = grouped && wgrad_tf32_prologue_ok(w);
  }
  std::vector<WgradJob> jobs1, jobs3;
  const bool sfused = skip_fused(p, d, c.B, c.mode);
  if (sfused) {
    // all R+1 skip wgrads in one launch: dw[n][(i, k)] = sum_p DO[p,n] * a_i[p,k]; the bias gradient (the same column
    // sums of DO for every skip conv) goes to the scratch and is added to each bias by the weight-norm backward
    RNVP_TRY(fork_to_side(c));
    ProfScope ps(PROF_WGRAD, S, 1, (R + 1) * d.D, d.D, c.wst);
    WgradArgs a{};
    a.x = c.act(ci, A.a[0]); a.dy = DO; a.dw = c.dw() + d.dwskip_off; a.dbias = c.dw() + d.dbskip_off;
    a.B = c.B; a.S = S; a.kpad = ld; a.n = d.D; a.npad = cv[1].npad; a.taps = 1; a.lddy = ld;
    a.segs = R + 1; a.seg_stride = A.a[1] - A.a[0]; a.lddw = (R + 1) * ld;
    a.x3 = x3_tier(p);
    RNVP_TRY(k_conv_wgrad_tf32(a, c.wst));
  }
  const float* DA = nullptr;               // d(a_{i+1}) accumulated so far
  for (int i = R - 1; i >= 0; --i) {
    const ConvDesc *rb0 = &cv[2 + 4 * i], *rb3 = rb0 + 1, *rb6 = rb0 + 2, *cs = rb0 + 3;
    float *ai = c.act(ci, A.a[i]), *an = c.act(ci, A.a[i + 1]);
    float *u1 = c.act(ci, A.u1[i]), *u2 = c.act(ci, A.u2[i]);
    // skip += core_skip_i(a_{i+1})
    float* DAc = take(3);
    if (!sfused) RNVP_TRY(run_wgrad(c, *cs, an, DO, ld, S, gbias(*cs)));
    RNVP_TRY(run_conv(c, *cs, true, DO, S, DAc, ld, nullptr, DA, nullptr, true));
    // a_{i+1} = a_i + rb6(relu(bn3(u2)))
    float* G3 = take(1);
    RNVP_TRY(recompute(3 * i + 2, u2));
    if (grouped) jobs1.push_back(WgradJob{u2, DAc, c.dw() + rb6->dw_off, gbias(*rb6), c.save(d.bns[3 * i + 2].save)});
    else RNVP_TRY(wgrad_bn(*rb6, 3 * i + 2, u2, DAc, ld, gbias(*rb6)));
    RNVP_TRY(dgrad_bn_bwd(*rb6, DAc, 3 * i + 2, G3, u2, G3, nullptr, true));
    // u2 = rb3(relu(bn2(u1)))
    float* G2 = take(2);
    RNVP_TRY(recompute(3 * i + 1, u1));
    if (grouped) jobs3.push_back(WgradJob{H, G3, c.dw() + rb3->dw_off, nullptr, nullptr});     // H = kept relu(bn2(u1))
    else RNVP_TRY(wgrad_bn(*rb3, 3 * i + 1, u1, G3, ld, nullptr));
    RNVP_TRY(dgrad_bn_bwd(*rb3, G3, 3 * i + 1, G2, u1, G2, nullptr, true));
    // u1 = rb0(relu(bn1(a_i)))
    float* G1 = take(1);
    float* DAn = take(3);
    RNVP_TRY(recompute(3 * i, ai));
    if (grouped) jobs1.push_back(WgradJob{ai, G2, c.dw() + rb0->dw_off, nullptr, c.save(d.bns[3 * i].save)});
    else RNVP_TRY(wgrad_bn(*rb0, 3 * i, ai, G2, ld, nullptr));
    RNVP_TRY(dgrad_bn_bwd(*rb0, G2, 3 * i, G1, ai, DAn, DAc, false));
    DA = DAn;
  }
  if (grouped) {
    RNVP_TRY(fork_to_side(c));               // every dy of the coupling's residual blocks has been produced
    const ConvDesc& r0 = cv[2], &r3 = cv[3];
    auto launch_group = [&](const std::vector<WgradJob>& jobs, const ConvDesc& cvw, bool xf) -> int {
      ProfScope ps(PROF_WGRAD, S, cvw.taps, (int)jobs.size() * cvw.cin, cvw.cout, c.wst);
      WgradArgs a{};
      a.B = c.B; a.S = S; a.kpad = cvw.kpad; a.n = cvw.cout; a.npad = cvw.npad; a.taps = cvw.taps; a.lddy = ld;
      a.lddw = cvw.ld_dw; a.x3 = x3_tier(p);
      a.xf_C = xf ? d.D : 0;
      a.njobs = (int)jobs.size(); a.jobs = jobs.data();
      if (a.njobs == 1) { a.x = jobs[0].x; a.dy = jobs[0].dy; a.dw = jobs[0].dw; a.dbias = jobs[0].dbias; a.xf_save = jobs[0].xf_save; }
      return k_conv_wgrad_tf32(a, c.wst);
    };
    RNVP_TRY(launch_group(jobs1, r0, true));
    RNVP_TRY(launch_group(jobs3, r3, false));
  }
  // skip = in_skip(a0) (+...); a0 = in_block(h0)
  float* DAf = take(3);
  if (!sfused) RNVP_TRY(run_wgrad(c, cv[1], c.act(ci, A.a[0]), DO, ld, S, gbias(cv[1])));
  RNVP_TRY(run_conv(c, cv[1], true, DO, S, DAf, ld, nullptr, DA, nullptr, true));
  RNVP_TRY(run_wgrad(c, cv[0], c.act(ci, A.h0), DAf, ld, S, gbias(cv[0])));
  RNVP_TRY(run_conv(c, cv[0], true, DAf, S, dh0, d.cin_pad, nullptr, nullptr, nullptr));
  if (fresh) RNVP_REQUIRE(next <= c.L.nbuf, "internal: backward scratch overrun (%d > %d)", next, c.L.nbuf);
  return RNVP_OK;
}

// ------------------------------------------------------------------------------------
// one coupling (NHWC in / out)
// ------------------------------------------------------------------------------------
int coupling_forward(const Ctx& c, int ci, const float* x, float* y, float* logJ, int training) {
  rnvp_plan* p = c.p;
  const CouplingDesc& d = p->cpl[ci];
  CplGeom g = d.geom(c.B);
  CplAct A = cpl_act(p, d, c.B, c.mode);
  const double count = (double)g.P() * p->world;
  if (training) {
    RNVP_TRY(k_cpl_in_stats(x, g, c.sf(d.sf_in), c.st));
    RNVP_TRY(sync_stats(c, c.sf(d.sf_in), 2 * d.cio));
  }
  RNVP_TRY(k_cpl_in_build(x, g, c.sf(d.sf_in), count, P_<float>(p, d, ci, SLOT_INBN_W),
                          P_<float>(p, d, ci, SLOT_INBN_B), P_<float>(p, d, ci, SLOT_INBN_RM),
                          P_<float>(p, d, ci, SLOT_INBN_RV), c.save(d.save_in), training, c.act(ci, A.h0),
                          p->math == RNVP_MATH_TF32, c.st));
  if (cpl_fused(p, d)) {
    // the affine map rides in the out conv's epilogue: training leaves x' + its batch statistics for out_bn (one
    // more tiny pass below); eval finishes the coupling there
    CplEpilogue e{};
    e.mode = training ? 1 : 2;
    e.store_st = training ? 1 : 0;
    e.g = g; e.x = x;
    e.out = training ? c.act(ci, A.xprime) : y;
    e.logJ = training ? nullptr : logJ;
    e.scale = P_<float>(p, d, ci, SLOT_SCALE); e.sshift = P_<float>(p, d, ci, SLOT_SSHIFT);
    e.run_mean = P_<float>(p, d, ci, SLOT_OUTBN_RM); e.run_var = P_<float>(p, d, ci, SLOT_OUTBN_RV);
    e.sums = c.sf(d.sf_out); e.logdet_acc = c.logdet_acc();
    RNVP_TRY(net_forward(c, ci, training, &e));
    if (!training) return RNVP_OK;
  } else {
    RNVP_TRY(net_forward(c, ci, training));
    RNVP_TRY(k_cpl_fwd_a(x, c.act(ci, A.st), g, P_<float>(p, d, ci, SLOT_SCALE), P_<float>(p, d, ci, SLOT_SSHIFT),
                         c.act(ci, A.xprime), c.sf(d.sf_out), c.logdet_acc(), training, c.st));
  }
  if (training) RNVP_TRY(sync_stats(c, c.sf(d.sf_out), 2 * d.cio));
  RNVP_TRY(k_cpl_fwd_b(c.act(ci, A.xprime), x, c.act(ci, A.st), g, c.sf(d.sf_out), count,
                       P_<float>(p, d, ci, SLOT_OUTBN_RM), P_<float>(p, d, ci, SLOT_OUTBN_RV),
                       c.save(d.save_out), training, P_<float>(p, d, ci, SLOT_SCALE),
                       P_<float>(p, d, ci, SLOT_SSHIFT), y, logJ, c.logdet_acc(), c.st));
  if (training) p->x_in[ci] = x;
  return RNVP_OK;
}

int coupling_inverse(const Ctx& c, int ci, const float* y, float* x, int training, float* logJ = nullptr) {
  rnvp_plan* p = c.p;
  const CouplingDesc& d = p->cpl[ci];
  CplGeom g = d.geom(c.B);
  CplAct A = cpl_act(p, d, c.B, c.mode);
  const double count = (double)g.P() * p->world;
  if (training) {
    RNVP_TRY(k_cpl_in_stats(y, g, c.sf(d.sf_in), c.st));
    RNVP_TRY(sync_stats(c, c.sf(d.sf_in), 2 * d.cio));
  }
  RNVP_TRY(k_cpl_in_build(y, g, c.sf(d.sf_in), count, P_<float>(p, d, ci, SLOT_INBN_W),
                          P_<float>(p, d, ci, SLOT_INBN_B), P_<float>(p, d, ci, SLOT_INBN_RM),
                          P_<float>(p, d, ci, SLOT_INBN_RV), c.save(d.save_in), training, c.act(ci, A.h0),
                          p->math == RNVP_MATH_TF32, c.st));
  if (cpl_fused(p, d)) {
    CplEpilogue e{};
    e.mode = 3;
    e.g = g; e.x = y; e.out = x; e.logJ = logJ;
    e.scale = P_<float>(p, d, ci, SLOT_SCALE); e.sshift = P_<float>(p, d, ci, SLOT_SSHIFT);
    e.run_mean = P_<float>(p, d, ci, SLOT_OUTBN_RM); e.run_var = P_<float>(p, d, ci, SLOT_OUTBN_RV);
    return net_forward(c, ci, training, &e);
  }
  RNVP_TRY(net_forward(c, ci, training));
  RNVP_TRY(k_cpl_inv(y, c.act(ci, A.st), g, P_<float>(p, d, ci, SLOT_OUTBN_RM), P_<float>(p, d, ci, SLOT_OUTBN_RV),
                     P_<float>(p, d, ci, SLOT_SCALE), P_<float>(p, d, ci, SLOT_SSHIFT), x, logJ, c.st));
  return RNVP_OK;
}

// dy -> dx (both NHWC [P,C]); dll (B) = dLoss/dlogdet per sample
int coupling_backward(const Ctx& c, int ci, const float* dy, const float* dll, float* dx) {
  rnvp_plan* p = c.p;
  const CouplingDesc& d = p->cpl[ci];
  CplGeom g = d.geom(c.B);
  CplAct A = cpl_act(p, d, c.B, c.mode);
  const double count = (double)g.P() * p->world;
  const float* x = p->x_in[ci];
  RNVP_REQUIRE(x != nullptr, "coupling %s: backward without a training forward", d.name.c_str());
  float* aux = c.aux(ci);
  size_t Pn = g.P();
  float* dst = aux;
  float* dxdir = dst + align_up(Pn * d.cst_pad, 64);
  float* dh0 = dxdir + align_up(Pn * d.cio, 64);
  const int set = c.L.sets == 2 ? (ci & 1) : 0;
  // the coupling that used this scratch set before must have finished on the side stream
  if (c.side_on && p->done_valid[set]) RNVP_CUDA(cudaStreamWaitEvent(c.st, p->ev_done[set], 0));
  RNVP_CUDA(cudaMemsetAsync(c.dw(), 0, d.dw_floats * 4, c.wst));
  RNVP_TRY(k_cpl_bwd_a(dy, c.act(ci, A.xprime), g, c.save(d.save_out), dll, c.sb(d.sb_cpl), c.st));
  RNVP_TRY(sync_stats(c, c.sb(d.sb_cpl), 2 * d.cio + 1));
  RNVP_TRY(k_cpl_bwd_b(dy, c.act(ci, A.xprime), x, c.act(ci, A.st), g, c.save(d.save_out), c.sb(d.sb_cpl), count,
                       dll, P_<float>(p, d, ci, SLOT_SCALE), P_<float>(p, d, ci, SLOT_SSHIFT), dst, dxdir,
                       G_(p, ci, SLOT_SCALE), G_(p, ci, SLOT_SSHIFT), p->math == RNVP_MATH_TF32, c.st));
  RNVP_TRY(net_backward(c, ci, dst, dh0));
  RNVP_TRY(k_cpl_in_bwd_a(dh0, x, g, c.save(d.save_in), c.sb(d.sb_in), c.st));
  RNVP_TRY(sync_stats(c, c.sb(d.sb_in), 2 * d.cio));
  RNVP_TRY(k_cpl_in_bwd_b(dh0, x, dxdir, dy, g, c.save(d.save_in), c.sb(d.sb_in), count,
                          P_<float>(p, d, ci, SLOT_INBN_W), G_(p, ci, SLOT_INBN_W), G_(p, ci, SLOT_INBN_B), dx,
                          1.0f / p->world, c.st));
  // weight-norm backward follows the wgrads of this coupling on their stream
  RNVP_TRY(k_weightnorm_bwd(p->d_jobs + d.job0, (int)d.convs.size(), p->max_cout, c.weights(), c.dw(), c.wst));
  if (c.side_on) {
    RNVP_CUDA(cudaEventRecord(p->ev_done[set], c.wst));
    p->done_valid[set] = true;
  }
  return RNVP_OK;
}

// first_cpl / ncpl: the couplings whose weights are (re)built (all of them for the flow, one for a stand-alone coupling)
int materialize_weights(const Ctx& c, int first_job, int njobs, int first_cpl, int ncpl) {
  RNVP_TRY(k_bias_sum(c.p->d_biasjobs + first_cpl, ncpl, c.weights(), c.st));
  return k_weightnorm_fwd(c.p->d_jobs + first_job, njobs, c.p->max_cout, c.weights(),
                          c.p->math == RNVP_MATH_TF32, c.st, x3_tier(c.p) ? c.p->weight_floats : 0);
}
// eval mode: the batch norms are fixed affine maps; their (scale, shift) go into the `save` arena once per call
int eval_bn_coefs(const Ctx& c) {
  return k_bn_eval_coefs(c.p->d_bnjobs, c.p->n_bnjobs, c.p->max_bn_c, c.save(0), c.st);
}

int zero_pass(const Ctx& c, bool backward) {
  if (!backward) {
    RNVP_CUDA(cudaMemsetAsync(c.ws + c.L.accum, 0, (32 + 2 * (size_t)c.B) * 8, c.st));
    if (c.L.stats_f_bytes) RNVP_CUDA(cudaMemsetAsync(c.ws + c.L.stats_f, 0, c.L.stats_f_bytes, c.st));
  } else if (c.L.stats_b_bytes) {
    RNVP_CUDA(cudaMemsetAsync(c.ws + c.L.stats_b, 0, c.L.stats_b_bytes, c.st));
  }
  return RNVP_OK;
}

// trunk tensors whose channel count is not a multiple of 32 carry zero padding that no kernel
// writes; clear the activation + scratch region once per call in that (test-sized) case
int clear_padding(const Ctx& c) {
  bool need = false;
  for (auto& d : c.p->cpl) need |= (d.ldD != d.D);
  if (need) RNVP_CUDA(cudaMemsetAsync(c.ws + c.L.act, 0, c.L.total - c.L.act, c.st));
  return RNVP_OK;
}

}  // namespace

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

const char* rnvp_last_error(void) { return rnvp::get_error(); }
const char* rnvp_version(void) { return "rnvp-b200 0.1 (sm_100a)"; }
unsigned long long rnvp_launch_count(void) { return rnvp::launch_count(); }

int rnvp_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(rnvp::g_prof_mu);
  rnvp::g_prof_on = on != 0;
  return RNVP_OK;
}
// Synchronises the device, aggregates the recorded intervals per kernel class and clears them.
// Each output row is (kind, S, taps, cin, cout, launches, total_ms); returns the number of rows
// written (<= max_rows) or a negative status.
int rnvp_prof_collect(double* rows, int max_rows) {
  RNVP_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(rnvp::g_prof_mu);
  std::map<std::vector<int>, std::pair<long, double>> agg;
  for (auto& r : rnvp::g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[{r.kind, r.S, r.taps, r.cin, r.cout}];
      e.first += 1;
      e.second += ms;
    }
    rnvp::g_ev_pool.push_back(r.a);
    rnvp::g_ev_pool.push_back(r.b);
  }
  rnvp::g_prof.clear();
  cudaGetLastError();
  int n = 0;
  for (auto& kv : agg) {
    if (n >= max_rows) break;
    double* o = rows + 7 * n++;
    for (int i = 0; i < 5; ++i) o[i] = kv.first[i];
    o[5] = (double)kv.second.first;
    o[6] = kv.second.second;
  }
  return n;
}

int rnvp_device_ok(void) {
  int dev = 0;
  RNVP_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  RNVP_CUDA(cudaGetDeviceProperties(&prop, dev));
  RNVP_REQUIRE(prop.major == 10, "device is sm_%d%d, this library is built for sm_100a only", prop.major, prop.minor);
  return RNVP_OK;
}

int rnvp_plan_create(const rnvp_config* cfg, rnvp_plan** out) {
  RNVP_REQUIRE(cfg && out, "null argument");
  rnvp_plan* p = new rnvp_plan();
  p->cfg = *cfg;
  int s = build_plan(p);
  if (s != RNVP_OK) {
    delete p;
    return s;
  }
  *out = p;
  return RNVP_OK;
}

int rnvp_plan_create_single(int kind, int C, int S, int D, int mask_cfg, int res_blocks, rnvp_plan** out) {
  RNVP_REQUIRE(out && (kind == 0 || kind == 1) && C > 0 && S > 0, "bad coupling description");
  rnvp_plan* p = new rnvp_plan();
  p->cfg = rnvp_config{C, S, D, res_blocks, 1, 0.f, 1.f};
  p->single = true;
  SingleSpec sp{kind, C, S, D, mask_cfg};
  int s = build_plan(p, &sp);
  if (s != RNVP_OK) {
    delete p;
    return s;
  }
  *out = p;
  return RNVP_OK;
}

int rnvp_plan_destroy(rnvp_plan* p) {
  if (!p) return RNVP_OK;
  if (p->d_jobs) cudaFree(p->d_jobs);
  if (p->d_segs) cudaFree(p->d_segs);
  if (p->d_bnjobs) cudaFree(p->d_bnjobs);
  if (p->d_biasjobs) cudaFree(p->d_biasjobs);
  if (p->side) {
    cudaStreamSynchronize(p->side);
    cudaStreamDestroy(p->side);
    for (auto& e : p->ev_ring) cudaEventDestroy(e);
    for (auto& e : p->ev_done) cudaEventDestroy(e);
    cudaEventDestroy(p->ev_join);
  }
  delete p;
  return RNVP_OK;
}

int rnvp_plan_num_couplings(const rnvp_plan* p) { return p ? (int)p->cpl.size() : 0; }
int rnvp_plan_slots_per_coupling(const rnvp_plan* p) { return p ? p->slots_per_coupling : 0; }
const char* rnvp_plan_slot_name(const rnvp_plan* p, int slot) {
  if (!p || slot < 0 || slot >= (int)p->slot_names.size()) return nullptr;
  return p->slot_names[slot].c_str();
}
int rnvp_plan_coupling_info(const rnvp_plan* p, int i, char* name, int name_len, int* kind, int* C, int* S,
                            int* D, int* mask_cfg) {
  RNVP_REQUIRE(p && i >= 0 && i < (int)p->cpl.size(), "coupling index out of range");
  const CouplingDesc& d = p->cpl[i];
  if (name && name_len > 0) snprintf(name, name_len, "%s", d.name.c_str());
  if (kind) *kind = d.kind;
  if (C) *C = d.C;
  if (S) *S = d.S;
  if (D) *D = d.D;
  if (mask_cfg) *mask_cfg = d.cfg;
  return RNVP_OK;
}

unsigned long long rnvp_plan_forward_generation(const rnvp_plan* p) { return p ? p->fwd_gen : 0; }

int rnvp_plan_set_math(rnvp_plan* p, int math) {
  RNVP_REQUIRE(p, "null plan");
  RNVP_REQUIRE(math == RNVP_MATH_FP32 || math == RNVP_MATH_TF32 || math == RNVP_MATH_TF32X3, "unknown math mode %d", math);
  p->math = math;
  return RNVP_OK;
}

int rnvp_plan_bind(rnvp_plan* p, void* const* params, void* const* grads, void* stream) {
  RNVP_REQUIRE(p && params, "null argument");
  size_t total = p->cpl.size() * (size_t)p->slots_per_coupling;
  for (size_t i = 0; i < total; ++i) {
    p->params[i] = params[i];
    p->grads[i] = grads ? grads[i] : nullptr;
  }
  std::vector<WnJob> jobs;
  std::vector<Seg> segs;
  for (size_t ci = 0; ci < p->cpl.size(); ++ci) {
    const CouplingDesc& d = p->cpl[ci];
    size_t base = ci * p->slots_per_coupling;
    for (int s = 0; s < p->slots_per_coupling; ++s) {
      if (p->slot_names[s].empty()) continue;        // bias slot of a bias-free conv
      RNVP_REQUIRE(p->params[base + s] != nullptr, "coupling %s: parameter '%s' is null", d.name.c_str(),
                   p->slot_names[s].c_str());
    }
    segs.push_back(Seg{(const float*)p->params[base + SLOT_SCALE], (float*)p->grads[base + SLOT_SCALE], 1});
    for (const ConvDesc& cv : d.convs) {
      WnJob j{};
      j.v = (const float*)p->params[base + cv.slot_v];
      j.g = (const float*)p->params[base + cv.slot_g];
      j.dv = (float*)p->grads[base + cv.slot_v];
      j.dg = cv.train_g ? (float*)p->grads[base + cv.slot_g] : nullptr;
      j.wf_off = cv.wf_off; j.wb_off = cv.wb_off; j.dw_off = cv.dw_off;
      j.cout = cv.cout; j.cin = cv.cin; j.taps = cv.taps;
      j.npad_f = cv.npad; j.kpad_f = cv.kpad; j.npad_b = cv.npad_b; j.kpad_b = cv.kpad_b;
      j.ld_f = cv.ld_f; j.ld_dw = cv.ld_dw;
      if (cv.is_skip && cv.has_bias) {       // a skip conv: its bias gradient may come from the fused wgrad
        j.dbias = (float*)p->grads[base + cv.slot_bias];
        j.dbias_src_off = d.dbskip_off;
      }
      jobs.push_back(j);
      if (cv.train_g) segs.push_back(Seg{j.g, (float*)p->grads[base + cv.slot_g], cv.cout});
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (!p->side) {
    int lo = 0, hi = 0;                         // lowest priority: the dgrad chain on the caller's stream goes first
    RNVP_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    RNVP_CUDA(cudaStreamCreateWithPriority(&p->side, cudaStreamNonBlocking, lo));
    for (auto& e : p->ev_ring) RNVP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : p->ev_done) RNVP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    RNVP_CUDA(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
  }
  std::vector<BiasSumJob> biasjobs;
  for (size_t ci = 0; ci < p->cpl.size(); ++ci) {
    const CouplingDesc& d = p->cpl[ci];
    size_t base = ci * p->slots_per_coupling;
    BiasSumJob bj{};
    for (const ConvDesc& cv : d.convs)
      if (cv.is_skip) bj.b[bj.n++] = (const float*)p->params[base + cv.slot_bias];
    bj.C = d.D; bj.out_off = d.bskip_off;
    biasjobs.push_back(bj);
  }
  if (!p->d_biasjobs) RNVP_CUDA(cudaMalloc(&p->d_biasjobs, biasjobs.size() * sizeof(BiasSumJob)));
  RNVP_CUDA(cudaMemcpyAsync(p->d_biasjobs, biasjobs.data(), biasjobs.size() * sizeof(BiasSumJob), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  std::vector<BnEvalJob> bnjobs;
  for (size_t ci = 0; ci < p->cpl.size(); ++ci) {
    const CouplingDesc& d = p->cpl[ci];
    size_t base = ci * p->slots_per_coupling;
    for (const BnDesc& b : d.bns) {
      bnjobs.push_back(BnEvalJob{(const float*)p->params[base + b.slot_w], (const float*)p->params[base + b.slot_b],
                                 (const float*)p->params[base + b.slot_rm], (const float*)p->params[base + b.slot_rv],
                                 b.save, b.C});
      p->max_bn_c = std::max(p->max_bn_c, b.C);
    }
  }
  if (!p->d_bnjobs) RNVP_CUDA(cudaMalloc(&p->d_bnjobs, bnjobs.size() * sizeof(BnEvalJob)));
  RNVP_CUDA(cudaMemcpyAsync(p->d_bnjobs, bnjobs.data(), bnjobs.size() * sizeof(BnEvalJob), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  p->n_bnjobs = (int)bnjobs.size();
  if (!p->d_jobs) RNVP_CUDA(cudaMalloc(&p->d_jobs, jobs.size() * sizeof(WnJob)));
  if (!p->d_segs) RNVP_CUDA(cudaMalloc(&p->d_segs, segs.size() * sizeof(Seg)));
  RNVP_CUDA(cudaMemcpyAsync(p->d_jobs, jobs.data(), jobs.size() * sizeof(WnJob), cudaMemcpyHostToDevice, st));
  RNVP_CUDA(cudaMemcpyAsync(p->d_segs, segs.data(), segs.size() * sizeof(Seg), cudaMemcpyHostToDevice, st));
  RNVP_CUDA(cudaStreamSynchronize(st));
  p->h_jobs = jobs;
  p->nsegs = (int)segs.size();
  p->bound = true;
  p->saved_batch = -1;
  p->saved_coupling = -1;
  return RNVP_OK;
}

size_t rnvp_plan_workspace_bytes(const rnvp_plan* p, int batch, int mode) {
  if (!p || batch <= 0) return 0;
  return compute_layout(p, batch, mode < 0 ? 0 : (mode > 2 ? 2 : mode)).total;
}

// ---- the flow ------------------------------------------------------------------------
}  // extern "C"

namespace {

// Flow-level buffers inside the workspace's flow region.  Group inputs and factored-out halves
// persist from a training forward to its backward.
struct FlowBufs {
  std::vector<float*> in_ckbd, in_chan, off;     // index by scale 1..L
  std::vector<size_t> n;                         // elements of the tensor entering scale s
  float* T[2];
  float* G[3];
  float* dll;
};
FlowBufs flow_bufs(const Ctx& c) {
  const int L = c.p->cfg.num_scales;
  FlowBufs f;
  f.in_ckbd.assign(L + 1, nullptr); f.in_chan.assign(L + 1, nullptr); f.off.assign(L + 1, nullptr);
  f.n.assign(L + 2, 0);
  float* base = reinterpret_cast<float*>(c.ws + c.L.flow);
  size_t o = 0;
  auto take = [&](size_t n) { float* r = base + o; o += align_up(n, 64); return r; };
  size_t n = c.L.img;
  for (int s = 1; s <= L; ++s) {
    f.n[s] = n;
    f.in_ckbd[s] = take(n);
    if (s < L) { f.in_chan[s] = take(n); f.off[s] = take(n / 2); }
    n /= 2;
  }
  f.T[0] = take(c.L.img); f.T[1] = take(c.L.img);
  f.G[0] = take(c.L.img); f.G[1] = take(c.L.img); f.G[2] = take(c.L.img);
  f.dll = take(c.B);
  return f;
}

}  // namespace

extern "C" {

int rnvp_flow_forward(rnvp_plan* p, const float* x_nchw, float* ll, float* logdet, float* z_nchw,
                      float* weight_scale, int batch, int training, void* ws, size_t ws_bytes, void* stream) {
  Ctx c;
  RNVP_REQUIRE(training >= 0 && training <= 2, "training must be 0, 1 or 2");
  RNVP_TRY(make_ctx(p, batch, training, ws, ws_bytes, stream, &c));
  RNVP_REQUIRE(!p->single, "flow entry points need a plan made by rnvp_plan_create");
  const rnvp_config& cf = p->cfg;
  const int L = cf.num_scales;
  p->saved_batch = -1;
  p->saved_coupling = -1;
  ++p->fwd_gen;
  RNVP_TRY(zero_pass(c, false));
  if (training && p->world > 1) RNVP_TRY(dp_check_equal_batches(&p->dp, batch, c.ws_acc() + 8, c.st));
  RNVP_TRY(clear_padding(c));
  RNVP_TRY(materialize_weights(c, 0, (int)p->h_jobs.size(), 0, (int)p->cpl.size()));
  if (!training) RNVP_TRY(eval_bn_coefs(c));
  if (weight_scale) {
    RNVP_TRY(k_sumsq(p->d_segs, p->nsegs, c.ws_acc(), c.st));
    RNVP_TRY(k_sumsq_finish(c.ws_acc(), weight_scale, c.st));
  }
  FlowBufs f = flow_bufs(c);
  RNVP_TRY(k_nchw_to_nhwc(x_nchw, f.in_ckbd[1], batch, cf.channels, cf.image_size, cf.image_size, c.st));
  const float* cur = f.in_ckbd[1];
  int ci = 0, chan = cf.channels, size = cf.image_size, flip = 0;
  auto run_group = [&](int n) -> int {
    for (int i = 0; i < n; ++i, ++ci) {
      CplAct A = cpl_act(p, p->cpl[ci], batch, c.mode);
      float* y = training ? c.act(ci, A.y) : f.T[flip ^= 1];
      RNVP_TRY(coupling_forward(c, ci, cur, y, nullptr, training));
      cur = y;
    }
    return RNVP_OK;
  };
  for (int s = 1; s < L; ++s) {
    RNVP_TRY(run_group(3));
    RNVP_TRY(k_permute(PERM_SQUEEZE, cur, nullptr, nullptr, nullptr, nullptr, f.in_chan[s], nullptr, nullptr, batch,
                       size / 2, chan, c.st));
    cur = f.in_chan[s];
    RNVP_TRY(run_group(3));
    // undo_squeeze o factor_out is a channel permutation at the squeezed resolution
    RNVP_TRY(k_permute(PERM_UNSQ_FACTOR, nullptr, cur, nullptr, nullptr, nullptr, nullptr, f.in_ckbd[s + 1], f.off[s],
                       batch, size / 2, chan, c.st));
    RNVP_TRY(k_prior_ll(f.off[s], batch, (int)(f.n[s] / 2 / batch), cf.prior_loc, cf.prior_scale, c.prior_acc(), c.st));
    cur = f.in_ckbd[s + 1];
    chan *= 2; size /= 2;
  }
  RNVP_TRY(run_group(4));
  RNVP_TRY(k_prior_ll(cur, batch, (int)(f.n[L] / batch), cf.prior_loc, cf.prior_scale, c.prior_acc(), c.st));
  RNVP_TRY(k_finalize_ll(c.logdet_acc(), c.prior_acc(), ll, logdet, batch, c.st));
  if (z_nchw) {
    // z = restore(...restore(z_L, off_{L-1})..., off_1)      flow_realnvp.py:316-325
    const float* t = cur;
    int ch = chan, sz = size, k = 0;
    for (int s = L - 1; s >= 1; --s) {
      float* dst = f.G[k ^= 1];
      RNVP_TRY(k_permute(PERM_RESTORE, nullptr, nullptr, t, f.off[s], dst, nullptr, nullptr, nullptr, batch, sz,
                         ch / 2, c.st));
      t = dst;
      ch /= 2; sz *= 2;
    }
    RNVP_TRY(k_nhwc_to_nchw(t, z_nchw, batch, cf.channels, cf.image_size, cf.image_size, c.st));
  }
  if (training) { p->saved_batch = batch; p->saved_mode = training; }
  return RNVP_OK;
}

int rnvp_flow_backward(rnvp_plan* p, const float* dll, const float* dweight_scale, float* dx_nchw, int batch, void* ws,
                       size_t ws_bytes, void* stream) {
  Ctx c;
  RNVP_TRY(make_ctx(p, batch, p->saved_mode, ws, ws_bytes, stream, &c));
  RNVP_REQUIRE(!p->single, "flow entry points need a plan made by rnvp_plan_create");
  if (p->saved_batch != batch) {
    set_error("rnvp_flow_backward: no matching training forward (saved batch %d, got %d)", p->saved_batch, batch);
    return RNVP_ERR_STATE;
  }
  const rnvp_config& cf = p->cfg;
  const int L = cf.num_scales;
  RNVP_TRY(zero_pass(c, true));
  if (dweight_scale != nullptr) RNVP_TRY(k_sumsq_bwd(p->d_segs, p->nsegs, dweight_scale, c.st));
  FlowBufs f = flow_bufs(c);
  int ci = (int)p->cpl.size() - 1;
  int chan = cf.channels << (L - 1), size = cf.image_size >> (L - 1), k = 0;
  // gradient of the prior term wrt the final latent block
  {
    const CouplingDesc& last = p->cpl[ci];
    CplAct A = cpl_act(p, last, batch, c.mode);
    RNVP_TRY(k_prior_grad(c.act(ci, A.y), dll, f.G[0], 0, batch, (int)(f.n[L] / batch), cf.prior_loc,
                          cf.prior_scale, c.st));
  }
  float* dcur = f.G[0];
  auto run_group = [&](int n) -> int {
    for (int i = 0; i < n; ++i, --ci) {
      float* dx = f.G[k ^= 1];
      RNVP_TRY(coupling_backward(c, ci, dcur, dll, dx));
      if (p->world > 1) {                      // overlap: reduce finished buckets
        RNVP_TRY(fork_to_side(c));             // the bucket needs both streams' gradients of this coupling
        RNVP_TRY(dp_coupling_done(&p->dp, ci, c.wst));
      }
      dcur = dx;
    }
    return RNVP_OK;
  };
  RNVP_TRY(run_group(4));
  for (int s = L - 1; s >= 1; --s) {
    chan /= 2; size *= 2;                     // (chan,size) of the tensor entering scale s
    // (on, off) = perm(sq): d_sq = perm^-1(d_on, d_off); d_off is the prior gradient of off_s
    RNVP_TRY(k_prior_grad(f.off[s], dll, f.G[2], 0, batch, (int)(f.n[s] / 2 / batch), cf.prior_loc, cf.prior_scale, c.st));
    float* dsq = f.G[k ^= 1];
    RNVP_TRY(k_permute(PERM_FACTOR_SQ, nullptr, nullptr, dcur, f.G[2], nullptr, dsq, nullptr, nullptr, batch,
                       size / 2, chan, c.st));
    dcur = dsq;
    RNVP_TRY(run_group(3));
    float* dhi = f.G[k ^= 1];
    RNVP_TRY(k_permute(PERM_UNDO_SQUEEZE, nullptr, dcur, nullptr, nullptr, dhi, nullptr, nullptr, nullptr, batch,
                       size / 2, chan, c.st));
    dcur = dhi;
    RNVP_TRY(run_group(3));
  }
  if (dx_nchw) RNVP_TRY(k_nhwc_to_nchw(dcur, dx_nchw, batch, cf.channels, cf.image_size, cf.image_size, c.st));
  RNVP_TRY(join_side(c));
  if (p->world > 1) RNVP_TRY(dp_join(&p->dp, c.st));
  p->saved_batch = -1;
  return RNVP_OK;
}

int rnvp_flow_inverse(rnvp_plan* p, const float* z_nchw, float* x_nchw, int batch, int training, void* ws,
                      size_t ws_bytes, void* stream) {
  Ctx c;
  RNVP_TRY(make_ctx(p, batch, 0, ws, ws_bytes, stream, &c));
  RNVP_REQUIRE(!p->single, "flow entry points need a plan made by rnvp_plan_create");
  const rnvp_config& cf = p->cfg;
  const int L = cf.num_scales;
  p->saved_batch = -1;
  p->saved_coupling = -1;
  ++p->fwd_gen;
  RNVP_TRY(zero_pass(c, false));
  RNVP_TRY(clear_padding(c));
  RNVP_TRY(materialize_weights(c, 0, (int)p->h_jobs.size(), 0, (int)p->cpl.size()));
  if (!training) RNVP_TRY(eval_bn_coefs(c));
  FlowBufs f = flow_bufs(c);
  // factor_out chain (flow_realnvp.py:197-200): in_ckbd[s+1], off[s] = factor_out(in_ckbd[s])
  RNVP_TRY(k_nchw_to_nhwc(z_nchw, f.in_ckbd[1], batch, cf.channels, cf.image_size, cf.image_size, c.st));
  int chan = cf.channels, size = cf.image_size;
  for (int s = 1; s < L; ++s) {
    RNVP_TRY(k_permute(PERM_FACTOR_OUT, f.in_ckbd[s], nullptr, nullptr, nullptr, nullptr, nullptr, f.in_ckbd[s + 1],
                       f.off[s], batch, size / 2, chan, c.st));
    chan *= 2; size /= 2;
  }
  const float* cur = f.in_ckbd[L];
  int ci = (int)p->cpl.size() - 1, flip = 0;
  auto run_group = [&](int n) -> int {
    for (int i = 0; i < n; ++i, --ci) {
      float* x = f.T[flip ^= 1];
      RNVP_TRY(coupling_inverse(c, ci, cur, x, training));
      cur = x;
    }
    return RNVP_OK;
  };
  RNVP_TRY(run_group(4));
  for (int s = L - 1; s >= 1; --s) {
    chan /= 2; size *= 2;
    // restore o squeeze: channel permutation at the low resolution
    float* sq = f.in_chan[s];
    RNVP_TRY(k_permute(PERM_FACTOR_SQ, nullptr, nullptr, cur, f.off[s], nullptr, sq, nullptr, nullptr, batch, size / 2,
                       chan, c.st));
    cur = sq;
    RNVP_TRY(run_group(3));
    float* hi = f.in_ckbd[s];
    RNVP_TRY(k_permute(PERM_UNDO_SQUEEZE, nullptr, cur, nullptr, nullptr, hi, nullptr, nullptr, nullptr, batch,
                       size / 2, chan, c.st));
    cur = hi;
    RNVP_TRY(run_group(3));
  }
  RNVP_TRY(k_nhwc_to_nchw(cur, x_nchw, batch, cf.channels, cf.image_size, cf.image_size, c.st));
  return RNVP_OK;
}

// ---- one coupling, NCHW boundary -------------------------------------------------------
int rnvp_coupling_forward(rnvp_plan* p, int ci, const float* x_nchw, float* y_nchw, float* logJ_nchw, int batch,
                          int training, void* ws, size_t ws_bytes, void* stream) {
  Ctx c;
  RNVP_REQUIRE(training >= 0 && training <= 2, "training must be 0, 1 or 2");
  RNVP_TRY(make_ctx(p, batch, training, ws, ws_bytes, stream, &c));
  RNVP_REQUIRE(ci >= 0 && ci < (int)p->cpl.size(), "coupling index %d out of range", ci);
  const CouplingDesc& d = p->cpl[ci];
  p->saved_batch = -1;
  p->saved_coupling = -1;
  ++p->fwd_gen;
  RNVP_TRY(zero_pass(c, false));
  RNVP_TRY(clear_padding(c));
  RNVP_TRY(materialize_weights(c, d.job0, (int)d.convs.size(), ci, 1));
  if (!training) RNVP_TRY(eval_bn_coefs(c));
  FlowBufs f = flow_bufs(c);
  RNVP_TRY(k_nchw_to_nhwc(x_nchw, f.T[0], batch, d.C, d.S, d.S, c.st));
  RNVP_TRY(coupling_forward(c, ci, f.T[0], f.T[1], logJ_nchw ? f.G[0] : nullptr, training));
  RNVP_TRY(k_nhwc_to_nchw(f.T[1], y_nchw, batch, d.C, d.S, d.S, c.st));
  if (logJ_nchw) RNVP_TRY(k_nhwc_to_nchw(f.G[0], logJ_nchw, batch, d.C, d.S, d.S, c.st));
  if (training) { p->saved_coupling = ci; p->saved_batch = batch; p->saved_mode = training; }
  return RNVP_OK;
}

int rnvp_coupling_inverse(rnvp_plan* p, int ci, const float* y_nchw, float* x_nchw, float* logJ_nchw, int batch,
                          int training, void* ws, size_t ws_bytes, void* stream) {
  Ctx c;
  RNVP_TRY(make_ctx(p, batch, 0, ws, ws_bytes, stream, &c));
  RNVP_REQUIRE(ci >= 0 && ci < (int)p->cpl.size(), "coupling index %d out of range", ci);
  const CouplingDesc& d = p->cpl[ci];
  p->saved_batch = -1;
  p->saved_coupling = -1;
  ++p->fwd_gen;
  RNVP_TRY(zero_pass(c, false));
  RNVP_TRY(clear_padding(c));
  RNVP_TRY(materialize_weights(c, d.job0, (int)d.convs.size(), ci, 1));
  if (!training) RNVP_TRY(eval_bn_coefs(c));
  FlowBufs f = flow_bufs(c);
  RNVP_TRY(k_nchw_to_nhwc(y_nchw, f.T[0], batch, d.C, d.S, d.S, c.st));
  RNVP_TRY(coupling_inverse(c, ci, f.T[0], f.T[1], training, logJ_nchw ? f.G[0] : nullptr));
  RNVP_TRY(k_nhwc_to_nchw(f.T[1], x_nchw, batch, d.C, d.S, d.S, c.st));
  if (logJ_nchw) RNVP_TRY(k_nhwc_to_nchw(f.G[0], logJ_nchw, batch, d.C, d.S, d.S, c.st));
  return RNVP_OK;
}

int rnvp_coupling_backward(rnvp_plan* p, int ci, const float* dy_nchw, const float* dlogJ_nchw, float* dx_nchw,
                           int batch, void* ws, size_t ws_bytes, void* stream) {
  Ctx c;
  RNVP_TRY(make_ctx(p, batch, p->saved_mode, ws, ws_bytes, stream, &c));
  if (p->saved_coupling != ci || p->saved_batch != batch) {
    set_error("rnvp_coupling_backward: no matching training forward of coupling %d at batch %d", ci, batch);
    return RNVP_ERR_STATE;
  }
  const CouplingDesc& d = p->cpl[ci];
  RNVP_TRY(zero_pass(c, true));
  FlowBufs f = flow_bufs(c);
  RNVP_TRY(k_nchw_to_nhwc(dy_nchw, f.G[0], batch, d.C, d.S, d.S, c.st));
  RNVP_TRY(k_gather_first(dlogJ_nchw, f.dll, batch, d.C * d.S * d.S, c.st));
  RNVP_TRY(coupling_backward(c, ci, f.G[0], f.dll, f.G[1]));
  RNVP_TRY(join_side(c));
  RNVP_TRY(k_nhwc_to_nchw(f.G[1], dx_nchw, batch, d.C, d.S, d.S, c.st));
  p->saved_coupling = -1;
  p->saved_batch = -1;
  return RNVP_OK;
}

// ---- logit, layout, building blocks ------------------------------------------------------
int rnvp_logit_forward(const float* x, const float* noise, float* y, float* logdet, int batch, int n,
                       float constraint, uint64_t seed, uint64_t offset, void* stream) {
  return k_logit_fwd(x, nullptr, noise, y, logdet, batch, n, constraint, seed, offset, (cudaStream_t)stream);
}
int rnvp_logit_forward_u8(const uint8_t* x, const float* noise, float* y, float* logdet, int batch, int n,
                          float constraint, uint64_t seed, uint64_t offset, void* stream) {
  return k_logit_fwd(nullptr, x, noise, y, logdet, batch, n, constraint, seed, offset, (cudaStream_t)stream);
}
int rnvp_logit_inverse(const float* y, float* x, size_t n, float constraint, void* stream) {
  return k_logit_inv(y, x, n, constraint, (cudaStream_t)stream);
}

}  // extern "C"

namespace {
// NCHW wrappers around the NHWC permute kernel need two scratch tensors; the layout entry points
// are API-completeness paths (the flow never calls them), so they allocate stream-ordered scratch.
struct Scratch {
  float* p = nullptr;
  cudaStream_t st;
  int alloc(size_t floats, cudaStream_t s) {
    st = s;
    RNVP_CUDA(cudaMallocAsync(&p, floats * 4, s));
    return RNVP_OK;
  }
  ~Scratch() { if (p) cudaFreeAsync(p, st); }
};
}  // namespace

extern "C" {

int rnvp_squeeze(const float* x, float* y, int B, int C, int H, int W, void* stream) {
  RNVP_REQUIRE(H == W && H % 2 == 0, "squeeze needs a square, even-sized input");
  cudaStream_t st = (cudaStream_t)stream;
  size_t n = (size_t)B * C * H * W;
  Scratch s;
  RNVP_TRY(s.alloc(2 * n, st));
  RNVP_TRY(k_nchw_to_nhwc(x, s.p, B, C, H, W, st));
  RNVP_TRY(k_permute(PERM_SQUEEZE, s.p, nullptr, nullptr, nullptr, nullptr, s.p + n, nullptr, nullptr, B, H / 2, C, st));
  return k_nhwc_to_nchw(s.p + n, y, B, 4 * C, H / 2, W / 2, st);
}
int rnvp_undo_squeeze(const float* x, float* y, int B, int C, int H, int W, void* stream) {
  RNVP_REQUIRE(H == W && C % 4 == 0, "undo_squeeze needs a square input with 4k channels");
  cudaStream_t st = (cudaStream_t)stream;
  size_t n = (size_t)B * C * H * W;
  Scratch s;
  RNVP_TRY(s.alloc(2 * n, st));
  RNVP_TRY(k_nchw_to_nhwc(x, s.p, B, C, H, W, st));
  RNVP_TRY(k_permute(PERM_UNDO_SQUEEZE, nullptr, s.p, nullptr, nullptr, s.p + n, nullptr, nullptr, nullptr, B, H, C / 4, st));
  return k_nhwc_to_nchw(s.p + n, y, B, C / 4, H * 2, W * 2, st);
}
int rnvp_factor_out(const float* x, float* on, float* off, int B, int C, int H, int W, void* stream) {
  RNVP_REQUIRE(H == W && H % 2 == 0, "factor_out needs a square, even-sized input");
  cudaStream_t st = (cudaStream_t)stream;
  size_t n = (size_t)B * C * H * W;
  Scratch s;
  RNVP_TRY(s.alloc(2 * n, st));
  RNVP_TRY(k_nchw_to_nhwc(x, s.p, B, C, H, W, st));
  RNVP_TRY(k_permute(PERM_FACTOR_OUT, s.p, nullptr, nullptr, nullptr, nullptr, nullptr, s.p + n, s.p + n + n / 2, B,
                     H / 2, C, st));
  RNVP_TRY(k_nhwc_to_nchw(s.p + n, on, B, 2 * C, H / 2, W / 2, st));
  return k_nhwc_to_nchw(s.p + n + n / 2, off, B, 2 * C, H / 2, W / 2, st);
}
int rnvp_restore(const float* on, const float* off, float* x, int B, int C, int H, int W, void* stream) {
  // on/off are (B,C,H,W); x is (B,C/2,2H,2W)
  RNVP_REQUIRE(H == W && C % 2 == 0, "restore needs square inputs with an even channel count");
  cudaStream_t st = (cudaStream_t)stream;
  size_t n = (size_t)B * C * H * W;
  Scratch s;
  RNVP_TRY(s.alloc(4 * n, st));
  RNVP_TRY(k_nchw_to_nhwc(on, s.p, B, C, H, W, st));
  RNVP_TRY(k_nchw_to_nhwc(off, s.p + n, B, C, H, W, st));
  RNVP_TRY(k_permute(PERM_RESTORE, nullptr, nullptr, s.p, s.p + n, s.p + 2 * n, nullptr, nullptr, nullptr, B, H, C / 2, st));
  return k_nhwc_to_nchw(s.p + 2 * n, x, B, C / 2, 2 * H, 2 * W, st);
}

int rnvp_weightnorm_forward(const float* v, const float* g, float* wf, float* wb, int cout, int cin, int ksize,
                            void* stream) {
  RNVP_REQUIRE(wf && wb, "both operand layouts are produced; pass two buffers");
  RNVP_REQUIRE(ksize == 1 || ksize == 3, "weight norm: kernel size %d unsupported (1 or 3)", ksize);
  cudaStream_t st = (cudaStream_t)stream;
  WnJob j{};
  j.v = v; j.g = g;
  j.cout = cout; j.cin = cin; j.taps = ksize * ksize;
  j.npad_f = pad_to(cout, 16); j.kpad_f = pad_to(cin, 32); j.npad_b = pad_to(cin, 16); j.kpad_b = pad_to(cout, 32);
  j.wf_off = 0;
  j.wb_off = (size_t)(wb - wf);               // relative to wf as the arena base
  RNVP_REQUIRE(wb > wf, "wb must follow wf in memory (single allocation)");
  Scratch s;
  RNVP_TRY(s.alloc(sizeof(WnJob) / 4 + 1, st));
  RNVP_CUDA(cudaMemcpyAsync(s.p, &j, sizeof(j), cudaMemcpyHostToDevice, st));
  RNVP_CUDA(cudaStreamSynchronize(st));
  return k_weightnorm_fwd(reinterpret_cast<WnJob*>(s.p), 1, cout, wf, 0, st);
}
int rnvp_weightnorm_backward(const float* v, const float* g, const float* dwf, float* dv, float* dg, int cout,
                             int cin, int ksize, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  WnJob j{};
  j.v = v; j.g = g; j.dv = dv; j.dg = dg;
  j.cout = cout; j.cin = cin; j.taps = ksize * ksize;
  j.npad_f = pad_to(cout, 16); j.kpad_f = pad_to(cin, 32); j.npad_b = pad_to(cin, 16); j.kpad_b = pad_to(cout, 32);
  j.dw_off = 0;
  Scratch s;
  RNVP_TRY(s.alloc(sizeof(WnJob) / 4 + 1, st));
  RNVP_CUDA(cudaMemcpyAsync(s.p, &j, sizeof(j), cudaMemcpyHostToDevice, st));
  RNVP_CUDA(cudaStreamSynchronize(st));
  return k_weightnorm_bwd(reinterpret_cast<WnJob*>(s.p), 1, cout, nullptr, dwf, st);
}

int rnvp_conv_forward(const float* x, const float* wf, const float* bias, const float* res, float* y, double* stats,
                      int B, int S, int kpad, int n, int npad, int ksize, int ldy, int math, void* stream) {
  ConvArgs a{};
  a.x = x; a.w = wf; a.bias = bias; a.res = res; a.y = y; a.stats = stats;
  a.B = B; a.S = S; a.kpad = kpad; a.n = n; a.npad = npad; a.taps = ksize * ksize; a.ldy = ldy;
  if (math == RNVP_MATH_TF32X3) {        // wf holds [w | w - trunc_tf32(w)], each taps * npad * kpad floats
    a.x3 = 1;
    a.w_lo_delta = (size_t)a.taps * npad * kpad;
  }
  return math != RNVP_MATH_FP32 ? k_conv_fwd_tf32(a, (cudaStream_t)stream) : k_conv_fwd_fp32(a, (cudaStream_t)stream);
}
int rnvp_conv_wgrad(const float* x, const float* dy, float* dwf, float* dbias, int B, int S, int kpad, int n,
                    int npad, int ksize, int lddy, int math, void* stream) {
  WgradArgs a{};
  a.x = x; a.dy = dy; a.dw = dwf; a.dbias = dbias;
  a.B = B; a.S = S; a.kpad = kpad; a.n = n; a.npad = npad; a.taps = ksize * ksize; a.lddy = lddy;
  a.x3 = math == RNVP_MATH_TF32X3;
  return math != RNVP_MATH_FP32 ? k_conv_wgrad_tf32(a, (cudaStream_t)stream) : k_conv_wgrad_fp32(a, (cudaStream_t)stream);
}

int rnvp_conv_forward_bn(const float* x_raw, const float* wf, const float* bias, const float* res, float* y,
                         double* stats, int B, int S, int kpad, int n, int npad, int ksize, int ldy, int bn_mode,
                         int bn_C, const double* bn_sums, double bn_count, const float* gamma, const float* beta,
                         float* run_mean, float* run_var, float* save, int round_out, void* stream) {
  ConvArgs a{};
  a.x = x_raw; a.w = wf; a.bias = bias; a.res = res; a.y = y; a.stats = stats;
  a.B = B; a.S = S; a.kpad = kpad; a.n = n; a.npad = npad; a.taps = ksize * ksize; a.ldy = ldy;
  a.round_out = round_out;
  RNVP_REQUIRE(conv_tf32_prologue_ok(a), "rnvp_conv_forward_bn: shape not supported by the tensor-core kernel");
  BnPrologue x{bn_mode, bn_C, bn_sums, bn_count, gamma, beta, run_mean, run_var, save};
  a.xf = &x;
  return k_conv_fwd_tf32(a, (cudaStream_t)stream);
}
int rnvp_conv_wgrad_bn(const float* x_raw, const float* dy, float* dwf, float* dbias, int B, int S, int kpad, int n,
                       int npad, int ksize, int lddy, const float* bn_save, int bn_C, void* stream) {
  WgradArgs a{};
  a.x = x_raw; a.dy = dy; a.dw = dwf; a.dbias = dbias;
  a.B = B; a.S = S; a.kpad = kpad; a.n = n; a.npad = npad; a.taps = ksize * ksize; a.lddy = lddy;
  a.xf_save = bn_save; a.xf_C = bn_C;
  RNVP_REQUIRE(bn_save != nullptr && wgrad_tf32_prologue_ok(a), "rnvp_conv_wgrad_bn: shape not supported by the tensor-core kernel");
  return k_conv_wgrad_tf32(a, (cudaStream_t)stream);
}

// ---- batch-norm building blocks (modules_realnvp.py:83-97, 139-143) on [P, ld] NHWC trunk tensors ----------
int rnvp_bn_relu_forward(const float* x, float* h, int P, int C, int ld, const double* sums, double count,
                         const float* gamma, const float* beta, float* run_mean, float* run_var, float* save,
                         int mode, int tf32_round, void* stream) {
  return k_bn_relu(x, h, P, C, ld, sums, count, gamma, beta, run_mean, run_var, save, mode, tf32_round,
                   (cudaStream_t)stream);
}
int rnvp_conv_dgrad_bn(const float* dy, const float* wb, const float* bn_x, const float* bn_save, float* gm,
                       double* sums2, int B, int S, int kpad, int n, int npad, int ksize, int ldy, void* stream) {
  ConvArgs a{};
  a.x = dy; a.w = wb; a.y = gm; a.stats = sums2; a.bn_x = bn_x; a.bn_save = bn_save;
  a.B = B; a.S = S; a.kpad = kpad; a.n = n; a.npad = npad; a.taps = ksize * ksize; a.ldy = ldy;
  RNVP_REQUIRE(conv_tf32_fusable(a), "rnvp_conv_dgrad_bn: shape not supported by the tensor-core kernel");
  return k_conv_fwd_tf32(a, (cudaStream_t)stream);
}
int rnvp_bn_backward_apply(const float* gm, const float* x, float* dx, const float* add, int P, int C, int ld,
                           const float* save, const double* sums2, double count, const float* gamma, float* dgamma,
                           float* dbeta, int raw_x_sums, int tf32_round, void* stream) {
  return k_bn_bwd_apply(gm, x, dx, add, P, C, ld, save, sums2, count, gamma, dgamma, dbeta, 1.0f, raw_x_sums,
                        tf32_round, (cudaStream_t)stream);
}

}  // extern "C"

namespace rnvp {
DpState* plan_dp(rnvp_plan* p) { return &p->dp; }
int plan_num_couplings(const rnvp_plan* p) { return (int)p->cpl.size(); }
}  // namespace rnvp
