// Hierarchy-specialised tree-head / tree-loss kernels, generated and compiled at run time.
//
// Why: the table-driven kernels (tree.cu, tree_lv.cu) interpret the hierarchy -- every node costs table loads, index
// arithmetic and shared-memory traffic -- and are instruction-issue bound far below the HBM roofline.  Here the
// hierarchy IS the program: for one nbdt_tree_desc this file emits a straight-line CUDA kernel (no loops over nodes, no
// tables, every constant an immediate), compiles it with NVRTC to an sm_100a cubin, caches the cubin on disk
// (lib/tree_cache/, pre-populated for the shipped hierarchies by build.py) and launches it through cudaLibrary*.
// Any hierarchy up to kMaxLeaves leaves gets such a kernel -- shipped, user-supplied or re-induced during training
// (nbdt/loss.py:303-315 in the reference) -- not only a build-time list.
//
// Kernel shape (one THREAD per row, rows of a warp are consecutive):
//   sweep A  the row is read once with the widest vector loads its size allows, max(x) is taken, and x is parked in
//            shared memory TRANSPOSED (slot c of thread t at word c*TPB + t: conflict-free scalar access, immediate
//            offsets).  4*C bytes of shared memory per row is the whole per-row footprint (+ spilled node values).
//   pass 1   post-order: subtree sums S[v] in registers; per inner node the child probabilities from the node logits
//            S[c]/#leaves(c) (binary node: one ex2 + one rcp; the SMALLER child probability is kept, its sign bit says
//            which child it belongs to, so both probabilities keep full relative accuracy); sum exp(x - max) for the
//            plain cross-entropy rides along.  Node values stay in REGISTERS up to a budget, the rest in shared memory.
//   pass 2   pre-order: path probabilities P[c] = P[parent]*p[c] (reference nbdt/model.py:207-242 multiplies
//            probabilities); at a leaf exp(P) by a degree-6 polynomial (FMA pipe), Z = sum exp(P), and the running
//            prefix F of E = exp(P)*P over the leaves in DFS order overwrites the (dead) x slot.
//   pass 4   pre-order: with leaves in DFS order every subtree sum of E is a difference of two prefix values and the
//            target's one-hot term is a step at the target's position, so the bottom-up gradient sweep disappears:
//            H[c] = T[end_c] - T[beg_c], dm = H[c] - p[c]*H[v], A[c] = A[v] + dm/#leaves(c); leaf slots receive A.
//   sweep B  x is read again (L2/L1 hit), dL/dx = A + w_x/B*(softmax(x) - onehot) is written with vector stores.
// The hard loss (loss.py:212-257) needs only pass 1 and ONE pre-order pass (its gradient lives on the target's path).
// HBM traffic = algorithmic bytes: logits + targets read once, gradient (or class probabilities) written once.
//
// Same closed form as tree.cu; reference semantics nbdt/model.py:83-123,207-242 and nbdt/loss.py:191-266.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "common.h"
#include "tree_jit.h"

namespace {

constexpr int kGeneratorVersion = 15;   // part of the cache key: bump when the emitted code changes
constexpr int kMaxLeaves = 1536;       // 32 rows x 132 B per leaf of shared memory per block: 1000 leaves (ImageNet) use 16 warps on 150 KB
constexpr int kTPB = 32;               // one warp per CTA: rows are thread-private, only __syncwarp around the cooperative I/O sweeps
// Shared-memory layout of a 32-row block: slot c (class c, later prefix / gradient) of row r lives at word
//   c*32 + c/U + rho(r),   rho(r) = (r*K) & 31,   U = elements per I/O vector,   K = ((C/U) & 31) | 1 (odd)
// Thread-per-row access (fixed c, r = lane): rho is a permutation of the banks -> conflict-free, and the offset of slot c
// from the thread's base (rho(lane) words) is an immediate.  Cooperative sweeps (a lane holds the U consecutive columns of
// one vector; consecutive lanes hold consecutive vectors, running on into the next row): for the k-th element of the
// vectors the bank is (vector index within the row + rho(row)) & 31, and rho advances by C/U (or C/U + 1) per row, so the 32
// lanes of an instruction hit 32 different banks (31 when C/U is even).  Spilled node values are thread-private: any bank
// permutation works, they follow the row slots.
constexpr int kRegValueCap = 168;      // node values kept in registers; the rest go to shared memory

// Tuning knobs (environment, read when a tree is created; non-default values become part of the cache key).
struct Knobs {
  int regcap = kRegValueCap;  // NBDT_B200_KA_REGCAP: node values kept in registers
  int minb = 0;               // NBDT_B200_KA_MINB: __launch_bounds__ min blocks (0 = derived from the register estimate)
  int unroll = 8;             // NBDT_B200_KA_UNROLL: unroll factor of the I/O sweeps (0 = full)
  int prefetch = 0;           // NBDT_B200_KA_PREFETCH: row blocks ahead whose lines are prefetched into L2 (0 = off; measured: no gain)
  int maxld = 24;             // NBDT_B200_KA_MAXLD: 16-byte loads a lane keeps in flight in sweep A before their stores
  int fence = -1;             // NBDT_B200_KA_FENCE: nodes between scheduling fences of the straight-line passes (0 = none; -1 = by
                              //   hierarchy size, measured in profiles/r02_tree_tune_v8.jsonl)
  int warps = 0;              // NBDT_B200_KA_WARPS: warps per row block (0 = by leaf count)
  int expm = -1;              // NBDT_B200_KA_EXPMUFU: exp(class probability) as one FMUL + MUFU.EX2 instead of the degree-6 polynomial
                              //   (6 FFMA): bit 0 in the Z traversal, bit 1 in the gradient traversal; -1 = by team size (measured,
                              //   profiles/r02_tree_tune_v6.jsonl: one warp per block 1, teams 3)
  int rows = 0;               // NBDT_B200_KA_ROWS: rows per thread (1, or 2 = packed f32x2 arithmetic; 0 = by leaf count)
  int teams = 0;              // NBDT_B200_KA_TEAMS: row blocks (teams) per CTA; the warps that run the SAME code for different row blocks
                              //   meet at a named barrier every `lockstep` nodes so that a fetched instruction line serves all of
                              //   them (0 = one team per CTA, no lock step)
  int lockstep = 16;          // NBDT_B200_KA_LOCKSTEP: nodes between those barriers
  bool is_default() const { return regcap == kRegValueCap && minb == 0 && unroll == 8 && prefetch == 0 && maxld == 24 && fence == -1 && warps == 0 && expm == -1 && rows == 0 && teams == 0 && lockstep == 16; }
};
Knobs read_knobs() {
  Knobs k;
  auto geti = [](const char* name, int dflt) { const char* e = getenv(name); return e && *e ? atoi(e) : dflt; };
  k.regcap = std::max(0, geti("NBDT_B200_KA_REGCAP", k.regcap));
  k.minb = std::max(0, std::min(32, geti("NBDT_B200_KA_MINB", k.minb)));
  k.unroll = std::max(0, geti("NBDT_B200_KA_UNROLL", k.unroll));
  k.prefetch = std::max(0, std::min(8, geti("NBDT_B200_KA_PREFETCH", k.prefetch)));
  k.maxld = std::max(1, std::min(64, geti("NBDT_B200_KA_MAXLD", k.maxld)));
  k.fence = std::max(-1, geti("NBDT_B200_KA_FENCE", k.fence));
  k.warps = std::max(0, std::min(16, geti("NBDT_B200_KA_WARPS", k.warps)));
  k.expm = std::max(-1, std::min(3, geti("NBDT_B200_KA_EXPMUFU", k.expm)));
  k.rows = std::max(0, std::min(2, geti("NBDT_B200_KA_ROWS", k.rows)));
  k.teams = std::max(0, std::min(8, geti("NBDT_B200_KA_TEAMS", k.teams)));
  k.lockstep = std::max(0, geti("NBDT_B200_KA_LOCKSTEP", k.lockstep));
  return k;
}

// ------------------------------------------------------------------------------------------------------------------
// hierarchy analysis
// ------------------------------------------------------------------------------------------------------------------
struct KTree {
  int C = 0, N = 0, NV = 0, root = 0;        // root: node id (C + index)
  std::vector<std::vector<int>> ch;          // [N] child node ids (leaf c -> c, inner i -> C + i)
  std::vector<int> cnt, beg, end;            // [NV] #leaves, DFS leaf range [beg, end)
  std::vector<int> pos, leaf_at;             // [C] DFS position of class c; class at DFS position
  unsigned long long hash = 0;
  int max_children = 0;
};

unsigned long long fnv(const nbdt_tree_desc* d) {
  unsigned long long h = 0xCBF29CE484222325ull;
  auto mix = [&](int v) {
    for (int b = 0; b < 4; ++b) {
      h ^= (unsigned long long)((v >> (8 * b)) & 0xFF);
      h *= 0x100000001B3ull;
    }
  };
  mix(d->num_classes); mix(d->num_inodes); mix(d->root);
  for (int i = 0; i <= d->num_inodes; ++i) mix(d->child_ptr[i]);
  for (int i = 0; i < d->child_ptr[d->num_inodes]; ++i) mix(d->child_node[i]);
  return h;
}

void analyse(const nbdt_tree_desc* d, KTree* t) {  // desc already validated by nbdt_tree_create
  t->C = d->num_classes; t->N = d->num_inodes; t->NV = t->C + t->N; t->root = t->C + d->root;
  t->ch.resize(t->N);
  for (int i = 0; i < t->N; ++i) {
    t->ch[i].assign(d->child_node + d->child_ptr[i], d->child_node + d->child_ptr[i + 1]);
    t->max_children = std::max(t->max_children, (int)t->ch[i].size());
  }
  t->cnt.assign(t->NV, 0); t->beg.assign(t->NV, 0); t->end.assign(t->NV, 0);
  t->pos.assign(t->C, 0); t->leaf_at.assign(t->C, 0);
  // iterative DFS in child order: leaves get consecutive positions
  struct Frame { int v; size_t k; };
  std::vector<Frame> st;
  st.push_back({t->root, 0});
  t->beg[t->root] = 0;
  int next = 0;
  while (!st.empty()) {
    Frame& f = st.back();
    const std::vector<int>& c = t->ch[f.v - t->C];
    if (f.k == c.size()) {
      t->end[f.v] = next;
      t->cnt[f.v] = next - t->beg[f.v];
      st.pop_back();
      continue;
    }
    const int u = c[f.k++];
    if (u < t->C) {
      t->pos[u] = next; t->leaf_at[next] = u; t->beg[u] = next; t->end[u] = next + 1; t->cnt[u] = 1;
      ++next;
    } else {
      t->beg[u] = next;
      st.push_back({u, 0});
    }
  }
  t->hash = fnv(d);
}

// ------------------------------------------------------------------------------------------------------------------
// code generation
// ------------------------------------------------------------------------------------------------------------------
struct Variant {
  int mode;    // kSpecSoftLoss / kSpecHardLoss / kSpecSoftFwd
  int dtype;   // NBDT_F32 / BF16 / F16
  int vec;     // bytes per logits/gradient vector access (16, 8, 4, 2)
  int ovec;    // bytes per class_probs vector store (soft forward)
  int layout = 0;  // 0: rows with stride ld.  1: class-major (N, C, H, W) logits / outputs, H*W a multiple of 32: the 32 rows of
                   // a block are 32 consecutive pixels of one image, every class is one coalesced line, and a thread fills its own
                   // slots (no cooperative transposition; the shared-memory layout needs no skew)
  int key() const { return (((mode * 4 + dtype) * 32 + vec) * 32 + ovec) * 2 + layout; }
};

// A team of W warps shares one block of 32 rows: the hierarchy is cut into disjoint subtrees ("cut subtrees", at most about
// C/(2W) leaves each) that are bin-packed onto the warps, and the few nodes above the cut ("top tree") go to warp 0.  Every
// warp runs the straight-line code of ITS subtrees for the same 32 rows (lane = row), so the node values a thread keeps in
// registers, the code a warp streams through the instruction cache and the shared memory per resident warp all shrink by W.
// The parts meet at bar.sync between the passes and exchange one value per cut root through shared memory.
struct Team {
  int W = 1;
  std::vector<int> cut_id;                     // [NV] index of the cut subtree rooted here, or -1
  std::vector<int> cut_nodes;                  // [ncut] root node of each cut subtree, in DFS order
  std::vector<int> part_of_cut;                // [ncut] warp that owns it
  std::vector<std::vector<int>> cuts_of_part;  // [W] cut indices, DFS order
  std::vector<int> part_of_node;               // [NV] owner warp (top-tree nodes: 0)
  std::vector<char> is_top;                    // [NV]
  int ntop = 0;
  int ncut() const { return (int)cut_nodes.size(); }
};

struct Plan {
  int nvalues = 0;                 // persistent node values (whole hierarchy)
  std::vector<int> value_slot;     // [NV]: first value id of node v's values (binary inner: 1 value; k-ary: one per child)
  std::vector<int> value_part;     // [nvalues] owner warp
  std::vector<int> value_spill;    // [nvalues] spill index within the owner's spill region, or -1 (register)
  int maxspill = 0;                // spill slots per warp
  int smem_bytes = 0;              // per CTA (upper bound over the variants' skew)
  int min_blocks = 1;              // __launch_bounds__ second argument
  int threads = kTPB;              // 32 * W
  int rows = 1;                    // rows per thread (2: packed pairs); a block is 32 * rows rows
  int teams = 1;                   // row blocks per CTA (CTA = teams x W warps, teams x the shared memory of one block)
  Team team;
  Knobs knobs;                     // with the automatic choices resolved
  bool default_knobs = true;       // nothing set in the environment: the variant name carries no knob suffix
  bool wide = false;               // soft loss without scheduling fences at <= 8 resident warps (see make_plan)
};

struct Out {
  std::string s;
  void f(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    va_list ap2;
    va_copy(ap2, ap);
    const int n = vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (n < (int)sizeof buf) {
      s.append(buf, (size_t)std::max(n, 0));
    } else {
      std::vector<char> big((size_t)n + 1);
      vsnprintf(big.data(), big.size(), fmt, ap2);
      s.append(big.data(), (size_t)n);
    }
    va_end(ap2);
    s.push_back('\n');
  }
};

std::string lit(double v) {  // exact fp32 literal
  char buf[64];
  snprintf(buf, sizeof buf, "%af", (double)(float)v);
  return buf;
}

constexpr double kLog2e = 1.4426950408889634;

const char* kPrelude = R"SRC(
// GENERATED by csrc/tree_jit.cu -- hierarchy-specialised straight-line tree kernel (see that file for the scheme).
struct KAParams {
  const void* x; long long B, ld; const long long* targets; float* class_probs; float* loss; void* grad;
  float w_xent, w_tree, inv_rows, grad_scale;
  long long hw;  // class-major variants: pixels per image
};
#ifndef KA_HOST
#define KA_DEV __device__ __forceinline__
typedef unsigned int u32;
typedef u32 KA_XSB;  // shared-memory byte address of the thread's slot 0
typedef unsigned short u16;
typedef unsigned long long u64;
KA_DEV float ex2(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
KA_DEV float lg2(float v) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
KA_DEV float rcp(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
KA_DEV float u2f(u32 v) { return __uint_as_float(v); }
KA_DEV u32 f2u(float v) { return __float_as_uint(v); }
KA_DEV float h2f(u32 v) { float r; asm("{ .reg .b16 h; cvt.u16.u32 h, %1; cvt.f32.f16 %0, h; }" : "=f"(r) : "r"(v)); return r; }
KA_DEV u32 pack_bf16(float lo, float hi) { u32 r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
KA_DEV u32 pack_f16(float lo, float hi) { u32 r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
KA_DEV uint4 ld16(const char* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
KA_DEV uint2 ld8(const char* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
KA_DEV u32 ld4(const char* p) { return __ldg(reinterpret_cast<const u32*>(p)); }
KA_DEV u32 ld2(const char* p) { return __ldg(reinterpret_cast<const u16*>(p)); }
KA_DEV void st16(char* p, u32 a, u32 b, u32 c, u32 d) { *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d); }
KA_DEV void st8(char* p, u32 a, u32 b) { *reinterpret_cast<uint2*>(p) = make_uint2(a, b); }
KA_DEV void st4(char* p, u32 a) { *reinterpret_cast<u32*>(p) = a; }
KA_DEV void st2(char* p, u32 a) { *reinterpret_cast<u16*>(p) = (u16)a; }
#endif

// ---- the value types of the row code: KF = one fp32 value per ROW OF THE THREAD ----
// KA_R == 1: a thread owns one row and KF is a float.  KA_R == 2: a thread owns two rows and KF is a packed pair in a 64-bit
// register: add / mul / fma are ONE instruction for both rows (add/mul/fma.rn.f32x2 -> FADD2 / FMUL2 / FFMA2 on sm_100; ptxas
// folds negation and |.| into their operand modifiers), MUFU ops, selects and sign tricks act on the halves.  The generated
// row code is the same text for both.
#if KA_R == 1
typedef float KF;
typedef bool KB;
typedef int KI;
typedef long long KY;
KA_DEV KF sel(KB c, KF a, KF b) { return c ? a : b; }
KA_DEV KF sel_neg(KF s, KF a, KF b) { return (int)f2u(s) < 0 ? a : b; }   // sign bit of s set
KA_DEV KF sel_ge0(KF t, KF a, KF b) { return t >= 0.f ? a : b; }
KA_DEV KF sel_eq(KI p, int k, KF a, KF b) { return p == k ? a : b; }
KA_DEV KF sel_lt(KI p, int k, KF a, KF b) { return p < k ? a : b; }
KA_DEV KF sel_in(KI p, int b0, int n, KF a, KF b) { return (unsigned)(p - b0) < (unsigned)n ? a : b; }
KA_DEV KB both(KB a, KB b) { return a && b; }
KA_DEV KB ka_true() { return true; }
KA_DEV bool ka_all(KB a) { return a; }
KA_DEV KB ka_not(KB a) { return !a; }
KA_DEV KF or_not_sign(KF v, KF t) { return u2f(f2u(v) | (~f2u(t) & 0x80000000u)); }
KA_DEV KF ka_fence(KF a, KF b, u32 zr) { return u2f(f2u(a) | (f2u(b) & zr)); }
KA_DEV u32 ka_bits(KF a) { return f2u(a); }
KA_DEV float ka_hsum(KF a) { return a; }
KA_DEV KB ka_valid(KY yy) { return (unsigned long long)yy < (unsigned long long)KA_C; }
KA_DEV KI ka_target(KY yy, KB valid) { return valid ? (int)(unsigned)yy : 0; }
KA_DEV KI ka_lookup(const unsigned short* tab, KI y) { return tab[y]; }
KA_DEV KB ka_owned(const unsigned char* tab, KI y, int w) { return tab[y] == w; }
#else
#ifdef KA_HOST
struct KF {
  float a, b;
  KF() {}
  KF(float f) : a(f), b(f) {}
  KF(float x, float y) : a(x), b(y) {}
};
KA_DEV float lo(KF v) { return v.a; }
KA_DEV float hi(KF v) { return v.b; }
KA_DEV KF operator+(KF x, KF y) { return KF(x.a + y.a, x.b + y.b); }
KA_DEV KF operator-(KF x, KF y) { return KF(x.a - y.a, x.b - y.b); }
KA_DEV KF operator*(KF x, KF y) { return KF(x.a * y.a, x.b * y.b); }
KA_DEV KF fmaf(KF x, KF y, KF z) { return KF(::fmaf(x.a, y.a, z.a), ::fmaf(x.b, y.b, z.b)); }
#else
struct KF {
  u64 v;
  KA_DEV KF() {}
  KA_DEV KF(float f) { asm("mov.b64 %0, {%1,%1};" : "=l"(v) : "f"(f)); }
  KA_DEV KF(float x, float y) { asm("mov.b64 %0, {%1,%2};" : "=l"(v) : "f"(x), "f"(y)); }
};
KA_DEV float lo(KF a) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return x; }
KA_DEV float hi(KF a) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return y; }
KA_DEV KF operator+(KF a, KF b) { KF r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
KA_DEV KF operator-(KF a, KF b) { KF r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
KA_DEV KF operator*(KF a, KF b) { KF r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
KA_DEV KF fmaf(KF a, KF b, KF c) { KF r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
#endif
KA_DEV KF operator-(KF a) { return KF(-lo(a), -hi(a)); }
KA_DEV KF& operator+=(KF& a, KF b) { a = a + b; return a; }
KA_DEV KF fmaxf(KF a, KF b) { return KF(::fmaxf(lo(a), lo(b)), ::fmaxf(hi(a), hi(b))); }
KA_DEV KF fabsf(KF a) { return KF(::fabsf(lo(a)), ::fabsf(hi(a))); }
KA_DEV KF ex2(KF a) { return KF(ex2(lo(a)), ex2(hi(a))); }
KA_DEV KF lg2(KF a) { return KF(lg2(lo(a)), lg2(hi(a))); }
KA_DEV KF rcp(KF a) { return KF(rcp(lo(a)), rcp(hi(a))); }
struct KB { bool a, b; };
struct KI { int a, b; };
struct KY { long long a, b; };
KA_DEV KB mkb(bool a, bool b) { KB r; r.a = a; r.b = b; return r; }
KA_DEV KI mki(int a, int b) { KI r; r.a = a; r.b = b; return r; }
KA_DEV KF sel(KB c, KF a, KF b) { return KF(c.a ? lo(a) : lo(b), c.b ? hi(a) : hi(b)); }
// selects of the straight-line passes, per half: setp + selp inside one asm statement so that ptxas emits ISETP + FSEL writing
// the half of the destination pair (a C++ ternary on the halves becomes two predicated MOVs per half)
#ifdef KA_HOST
KA_DEV float ka_sneg(float s, float a, float b) { return (int)f2u(s) < 0 ? a : b; }
KA_DEV float ka_sge0(float t, float a, float b) { return t >= 0.f ? a : b; }
KA_DEV float ka_seq(int p, int k, float a, float b) { return p == k ? a : b; }
KA_DEV float ka_slt(int p, int k, float a, float b) { return p < k ? a : b; }
KA_DEV float ka_sltu(u32 p, u32 k, float a, float b) { return p < k ? a : b; }
#else
KA_DEV float ka_sneg(float s, float a, float b) { float r; asm("{ .reg .pred p; setp.lt.s32 p, %3, 0; selp.f32 %0, %1, %2, p; }" : "=f"(r) : "f"(a), "f"(b), "r"(f2u(s))); return r; }
KA_DEV float ka_sge0(float t, float a, float b) { float r; asm("{ .reg .pred p; setp.ge.f32 p, %3, 0f00000000; selp.f32 %0, %1, %2, p; }" : "=f"(r) : "f"(a), "f"(b), "f"(t)); return r; }
KA_DEV float ka_seq(int p_, int k, float a, float b) { float r; asm("{ .reg .pred p; setp.eq.s32 p, %3, %4; selp.f32 %0, %1, %2, p; }" : "=f"(r) : "f"(a), "f"(b), "r"(p_), "r"(k)); return r; }
KA_DEV float ka_slt(int p_, int k, float a, float b) { float r; asm("{ .reg .pred p; setp.lt.s32 p, %3, %4; selp.f32 %0, %1, %2, p; }" : "=f"(r) : "f"(a), "f"(b), "r"(p_), "r"(k)); return r; }
KA_DEV float ka_sltu(u32 p_, u32 k, float a, float b) { float r; asm("{ .reg .pred p; setp.lt.u32 p, %3, %4; selp.f32 %0, %1, %2, p; }" : "=f"(r) : "f"(a), "f"(b), "r"(p_), "r"(k)); return r; }
#endif
KA_DEV KF sel_neg(KF s, KF a, KF b) { return KF(ka_sneg(lo(s), lo(a), lo(b)), ka_sneg(hi(s), hi(a), hi(b))); }
KA_DEV KF sel_ge0(KF t, KF a, KF b) { return KF(ka_sge0(lo(t), lo(a), lo(b)), ka_sge0(hi(t), hi(a), hi(b))); }
KA_DEV KF sel_eq(KI p, int k, KF a, KF b) { return KF(ka_seq(p.a, k, lo(a), lo(b)), ka_seq(p.b, k, hi(a), hi(b))); }
KA_DEV KF sel_lt(KI p, int k, KF a, KF b) { return KF(ka_slt(p.a, k, lo(a), lo(b)), ka_slt(p.b, k, hi(a), hi(b))); }
KA_DEV KF sel_in(KI p, int b0, int n, KF a, KF b) { return KF(ka_sltu((u32)(p.a - b0), (u32)n, lo(a), lo(b)), ka_sltu((u32)(p.b - b0), (u32)n, hi(a), hi(b))); }
KA_DEV KB both(KB a, KB b) { return mkb(a.a && b.a, a.b && b.b); }
KA_DEV KB ka_true() { return mkb(true, true); }
KA_DEV bool ka_all(KB a) { return a.a && a.b; }
KA_DEV KB ka_not(KB a) { return mkb(!a.a, !a.b); }
KA_DEV KF or_not_sign(KF v, KF t) { return KF(u2f(f2u(lo(v)) | (~f2u(lo(t)) & 0x80000000u)), u2f(f2u(hi(v)) | (~f2u(hi(t)) & 0x80000000u))); }
// one half carries the dependency: the packed instructions that consume the pair are ordered with it
KA_DEV KF ka_fence(KF a, KF b, u32 zr) { return KF(u2f(f2u(lo(a)) | (f2u(lo(b)) & zr)), hi(a)); }
KA_DEV u32 ka_bits(KF a) { return f2u(lo(a)); }
KA_DEV float ka_hsum(KF a) { return lo(a) + hi(a); }
KA_DEV KB ka_valid(KY yy) { return mkb((unsigned long long)yy.a < (unsigned long long)KA_C, (unsigned long long)yy.b < (unsigned long long)KA_C); }
KA_DEV KI ka_target(KY yy, KB valid) { return mki(valid.a ? (int)(unsigned)yy.a : 0, valid.b ? (int)(unsigned)yy.b : 0); }
KA_DEV KI ka_lookup(const unsigned short* tab, KI y) { return mki(tab[y.a], tab[y.b]); }
KA_DEV KB ka_owned(const unsigned char* tab, KI y, int w) { return mkb(tab[y.a] == w, tab[y.b] == w); }
#endif

#ifdef KA_HOST
// host harness (tests/jit_host.py): the "shared memory" of a block is a float array, slot s of the thread's KA_R rows at s*KA_R
#if KA_R == 1
#define KA_LDX(dst, base, slot, off) const KF dst = (base)[slot]
#define KA_STX(base, slot, off, val) (base)[slot] = (val)
#define KA_LDX_DYN(dst, idx) const KF dst = xsb[idx]
#define KA_STX_DYN(idx, val) xsb[idx] = (val)
#define KA_ADDX_DYN(cond, idx, val) if (cond) xsb[idx] += (val)
#define KA_ZERO_IF(cond, idx) if (cond) xsb[idx] = 0.f
#else
#define KA_LDX(dst, base, slot, off) const KF dst = KF((base)[(slot) * 2], (base)[(slot) * 2 + 1])
#define KA_STX(base, slot, off, val) { const KF sv_ = (val); (base)[(slot) * 2] = lo(sv_); (base)[(slot) * 2 + 1] = hi(sv_); }
#define KA_LDX_DYN(dst, idx) const KF dst = KF(xsb[(idx).a * 2], xsb[(idx).b * 2 + 1])
#define KA_ADDX_DYN(cond, idx, val) { const KB dc_ = (cond); if (dc_.a) xsb[(idx).a * 2] += (val); if (dc_.b) xsb[(idx).b * 2 + 1] += (val); }
#define KA_ZERO_IF(cond, idx) { const KB dc_ = (cond); if (dc_.a) xsb[(idx) * 2] = 0.f; if (dc_.b) xsb[(idx) * 2 + 1] = 0.f; }
#endif
#define KA_OPAQUE(x)
#define KA_SLOT_ADDR(base, col, row) ((base) + (col) * KA_R + (row))
#define KA_STD(a, k, off, val) (a)[(k) * KA_R] = (val)
#define KA_LDD(dst, a, k, off) const float dst = (a)[(k) * KA_R]
#else
// Shared-memory accesses are volatile inline PTX with immediate offsets: plain C++ accesses get forwarded by the compiler into
// registers and from there into local-memory spills.  Slot `idx` of the thread: KA_R floats at byte KA_DYN_OFF(idx) from xsb.
#define KA_DYN_OFF(idx) (((u32)(idx) * 128u + ((u32)(idx) / KA_U) * 4u) * KA_R)
#if KA_R == 1
#define KA_LDX(dst, base, slot, off) KF dst; asm volatile("ld.shared.f32 %0, [%1+" #off "];" : "=f"(dst) : "r"(base))
#define KA_STX(base, slot, off, val) asm volatile("st.shared.f32 [%0+" #off "], %1;" :: "r"(base), "f"((KF)(val)))
#define KA_OPAQUE(x) asm volatile("" : "+f"(x))
#define KA_LDX_DYN(dst, idx) KF dst; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(dst) : "r"(xsb + KA_DYN_OFF(idx)))
#define KA_STX_DYN(idx, val) asm volatile("st.shared.f32 [%0], %1;" :: "r"(xsb + KA_DYN_OFF(idx)), "f"((KF)(val)))
#define KA_ADDX_DYN(cond, idx, val) if (cond) { float dv_; const u32 da_ = xsb + KA_DYN_OFF(idx); asm volatile("ld.shared.f32 %0, [%1];" : "=f"(dv_) : "r"(da_)); dv_ += (val); asm volatile("st.shared.f32 [%0], %1;" :: "r"(da_), "f"(dv_)); }
#define KA_ZERO_IF(cond, idx) if (cond) asm volatile("st.shared.f32 [%0], %1;" :: "r"(xsb + KA_DYN_OFF(idx)), "f"(0.f))
#else
#define KA_LDX(dst, base, slot, off) KF dst; asm volatile("ld.shared.b64 %0, [%1+" #off "];" : "=l"(dst.v) : "r"(base))
#define KA_STX(base, slot, off, val) { const KF sv_ = (val); asm volatile("st.shared.b64 [%0+" #off "], %1;" :: "r"(base), "l"(sv_.v)); }
#define KA_OPAQUE(x) asm volatile("" : "+l"((x).v))
KA_DEV float ka_lds(u32 a) { float r; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a)); return r; }
KA_DEV void ka_sts(u32 a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v)); }
#define KA_LDX_DYN(dst, idx) const KF dst = KF(ka_lds(xsb + KA_DYN_OFF((idx).a)), ka_lds(xsb + KA_DYN_OFF((idx).b) + 4u))
#define KA_ADDX_DYN(cond, idx, val) { const KB dc_ = (cond); \
    if (dc_.a) { const u32 da_ = xsb + KA_DYN_OFF((idx).a); ka_sts(da_, ka_lds(da_) + (val)); } \
    if (dc_.b) { const u32 da_ = xsb + KA_DYN_OFF((idx).b) + 4u; ka_sts(da_, ka_lds(da_) + (val)); } }
#define KA_ZERO_IF(cond, idx) { const KB dc_ = (cond); if (dc_.a) ka_sts(xsb + KA_DYN_OFF(idx), 0.f); if (dc_.b) ka_sts(xsb + KA_DYN_OFF(idx) + 4u, 0.f); }
#endif
// cooperative sweeps: slot `col` (a multiple of the vector size) of row `row` of the block (thread row / KA_R, half row % KA_R),
// and accesses relative to it
#define KA_SLOT_ADDR(base, col, row) ((base) + (((u32)(col) * 32u + (u32)(col) / KA_U + ((((u32)(row) / KA_R) * KA_K) & 31u)) * KA_R + (u32)(row) % KA_R) * 4u)
#define KA_STD(a, k, off, val) asm volatile("st.shared.f32 [%0+" #off "], %1;" :: "r"(a), "f"(val))
#define KA_LDD(dst, a, k, off) float dst; asm volatile("ld.shared.f32 %0, [%1+" #off "];" : "=f"(dst) : "r"(a))
// the parts of a team meet here (all threads of the CTA); a single-warp kernel needs no barrier inside the row function
#if KA_T > 1
// several teams per CTA: a team's barrier is named barrier 1 + team (32 * KA_W threads); KA_LOCKSTEP is the named barrier of the
// KA_T warps that run the same part for different row blocks
#if KA_W > 1
#define KA_BAR() asm volatile("bar.sync %0, %1;" :: "r"(1u + threadIdx.x / (32u * KA_W)), "n"(32 * KA_W) : "memory")
#else
#define KA_BAR()
#endif
#define KA_LOCKSTEP() asm volatile("bar.sync %0, %1;" :: "r"(1u + KA_T + ((threadIdx.x >> 5) % KA_W)), "n"(32 * KA_T) : "memory")
#elif KA_W > 1
#define KA_BAR() asm volatile("bar.sync 0;" ::: "memory")
#define KA_LOCKSTEP()
#else
#define KA_BAR()
#define KA_LOCKSTEP()
#endif
#endif
#ifdef KA_HOST
#define KA_LOCKSTEP()
#endif
// exp(t), t in [0, 1] (a class probability): degree-6 interpolant on the FMA pipe, max rel. error 1.5e-7
KA_DEV KF exp01(KF t) {
  KF r = 0.002307842718437314f;
  r = fmaf(r, t, 0.0069595566019415855f);
  r = fmaf(r, t, 0.04264117032289505f);
  r = fmaf(r, t, 0.16631804406642914f);
  r = fmaf(r, t, 0.5000588297843933f);
  r = fmaf(r, t, 0.9999963045120239f);
  return fmaf(r, t, 1.0f);
}
KA_DEV KF exp01m(KF t) { return ex2(t * 1.4426950408889634f); }  // the same on the MUFU pipe (2 ulp)
// smaller child probability of a binary node with the sign bit set when it belongs to the SECOND child
// (t = node logit of the first child minus that of the second, in log2 units)
KA_DEV KF signed_lo(KF t) {
  const KF q = ex2(-fabsf(t));
  return or_not_sign(q * rcp(1.f + q), t);
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
)SRC";

struct Gen {
  const KTree& t;
  const Variant v;
  const Plan& plan;
  const Team& team;
  Out o;
  int acc = 0;       // round-robin partial accumulators
  int part = 0;      // warp whose code is being emitted
  bool top = false;  // emitting the top tree (stops at the cut roots)
  Gen(const KTree& t_, const Variant& v_, const Plan& p_) : t(t_), v(v_), plan(p_), team(p_.team) {}

  bool leaf(int n) const { return n < t.C; }
  const std::vector<int>& kids(int n) const { return t.ch[n - t.C]; }
  bool binary(int n) const { return kids(n).size() == 2; }
  bool loss_mode() const { return v.mode != kSpecSoftFwd; }
  int esz() const { return v.dtype == NBDT_F32 ? 4 : 2; }
  bool teamed() const { return team.W > 1; }
  bool is_cut(int n) const { return teamed() && team.cut_id[n] >= 0; }

  // ---- scheduling fences ----
  // The passes are one basic block of thousands of instructions with hundreds of independent chains; left alone ptxas
  // front-loads as many of them as it can and spills hundreds of registers.  Every `fence` nodes the next node's input is
  // OR-ed with (last value produced & zr), zr a run-time zero: a real data dependency that changes no value and costs one
  // LOP3, so the scheduler's window is a group of nodes rather than the whole pass.
  int tmp = 0, since = 0, nfence = 0;
  std::string last;
  std::string base = "xsb";
  int lock_since = 0;
  void note(const std::string& name) {
    last = name; ++since;
    if (plan.teams > 1 && plan.knobs.lockstep > 0 && ++lock_since >= plan.knobs.lockstep) {
      lock_since = 0;
      o.f("    KA_LOCKSTEP();");
    }
  }
  // the no-fence / 8-warp setting was measured on the soft loss only (its two traversals carry the long dependency chains);
  // the hard loss and the soft forward keep the fenced code at full occupancy (hard loss, 100 leaves: 58 % vs 51 %)
  bool wide_mode() const { return plan.wide && v.mode == kSpecSoftLoss; }
  int fence_every() const { return plan.wide && !wide_mode() ? 8 : plan.knobs.fence; }
  int min_blocks() const { return wide_mode() ? std::min(plan.min_blocks, std::max(1, 8 / plan.teams)) : plan.min_blocks; }
  bool fence_due() const { return fence_every() > 0 && since >= fence_every() && !last.empty(); }
  std::string fenced(const std::string& val) {
    if (!fence_due()) return val;
    since = 0;
    char b[32];
    snprintf(b, sizeof b, "fz%d", nfence++);
    o.f("    const KF %s = ka_fence(%s, %s, zr);", b, val.c_str(), last.c_str());
    return b;
  }
  void fence_base() {  // same for the shared-memory base of the following loads / stores
    if (!fence_due()) return;
    since = 0;
    char b[32];
    snprintf(b, sizeof b, "xb%d", nfence++);
    o.f("    const KA_XSB %s = xsb + (ka_bits(%s) & zr);", b, last.c_str());
    base = b;
  }

  // ---- shared-memory slots ----
  // slots 0..C-1: the row (x, later dL/dx or the class probabilities); then W spill regions of plan.maxspill slots; then the
  // exchange slots of the team.  Byte offsets are relative to the thread's base (rho(lane) words into the block).
  int U = 1, K = 1;  // skew unit and row multiplier of the layout (see the layout comment at the top of this file)
  int R() const { return plan.rows; }  // rows per thread: every slot holds R consecutive floats
  int row_words() const { return (t.C * 32 + t.C / U + 32 + 3) & ~3; }
  int soff(int slot) const {
    if (slot < t.C) return (slot * 32 + slot / U) * 4 * R();
    return (row_words() + (slot - t.C) * 32) * 4 * R();
  }
  int eoff(int k) const { return (k * 32 + k / U) * 4 * R(); }  // element k of a vector, relative to the vector's first slot
  int spill_slot(int vid) const { return t.C + plan.value_part[vid] * plan.maxspill + plan.value_spill[vid]; }
  int ex_base() const { return t.C + team.W * plan.maxspill; }
  int ex_p(int k) const { return ex_base() + k; }                      // subtree sum S, then path probability P (hard: on-path flag)
  int ex_h(int k) const { return ex_base() + team.ncut() + k; }        // gradient sum H, then accumulated gradient A
  int ex_w(int j, int w) const { return ex_base() + 2 * team.ncut() + j * team.W + w; }  // per-warp partials: 0 max, 1 sum exp, 2 Z, 3 P[target] / path loss
  int total_slots() const { return ex_base() + 2 * team.ncut() + 4 * team.W; }
  // Emitted as volatile inline PTX with immediate offsets: plain C++ accesses get forwarded by the compiler into registers and
  // from there into local-memory spills.
  std::string ldx(int slot) {
    char b[32];
    snprintf(b, sizeof b, "t%d", tmp++);
    o.f("    KA_LDX(%s, %s, %d, %d);", b, base.c_str(), slot, soff(slot));
    return b;
  }
  void stx(int slot, const std::string& val) { o.f("    KA_STX(%s, %d, %d, %s);", base.c_str(), slot, soff(slot), val.c_str()); }

  // ---- persistent node values: register variable or spill slot of the owning warp ----
  // value id of a binary node = plan.value_slot[node]; of child k of a k-ary node = plan.value_slot[node] + k
  bool spilled(int vid) const { return plan.value_spill[vid] >= 0; }
  std::string vget(int vid) {
    if (spilled(vid)) return ldx(spill_slot(vid));
    return "nv" + std::to_string(vid);
  }
  void vset(int vid, const std::string& expr) {
    if (spilled(vid)) stx(spill_slot(vid), expr);
    else o.f("    const KF nv%d = %s;", vid, expr.c_str());
  }
  std::string kget(int vid) {  // k-ary child value
    if (spilled(vid)) return ldx(spill_slot(vid));
    return "km" + std::to_string(vid);
  }
  // gradient values: same placement as the probability values they replace
  void dset(int vid, const std::string& expr) {
    if (spilled(vid)) stx(spill_slot(vid), expr);
    else o.f("    const KF dv%d = %s;", vid, expr.c_str());
  }
  std::string dget(int vid) {
    if (spilled(vid)) return ldx(spill_slot(vid));
    return "dv" + std::to_string(vid);
  }

  // ---- row vectors ----
  int per_vec() const { return v.vec / esz(); }
  std::string word(const char* name, int i, int w) const {  // w-th 32-bit word of vector i
    char b[48];
    static const char* comp[4] = {"x", "y", "z", "w"};
    if (v.vec >= 8) snprintf(b, sizeof b, "%s%d.%s", name, i, comp[w]);
    else snprintf(b, sizeof b, "%s%d", name, i);
    return b;
  }
  std::string elem(const char* name, int c) const {  // fp32 value of element c out of the loaded vectors name0, name1, ...
    const int pv = per_vec(), i = c / pv, j = c % pv;
    char b[96];
    if (v.dtype == NBDT_F32) {
      snprintf(b, sizeof b, "u2f(%s)", word(name, i, j).c_str());
    } else {
      const std::string w = word(name, i, j / 2);
      if (v.dtype == NBDT_BF16) {
        if (v.vec == 2 || j % 2 == 0) snprintf(b, sizeof b, "u2f(%s << 16)", w.c_str());
        else snprintf(b, sizeof b, "u2f(%s & 0xffff0000u)", w.c_str());
      } else {
        if (v.vec == 2 || j % 2 == 0) snprintf(b, sizeof b, "h2f(%s)", w.c_str());
        else snprintf(b, sizeof b, "h2f(%s >> 16)", w.c_str());
      }
    }
    return b;
  }

  // ---- pass 1: post-order.  Subtree sums in registers; per inner node the persistent value ----
  void up(int n) {
    if (top && is_cut(n)) {  // top tree: the sum of a cut subtree comes from its owner
      fence_base();
      o.f("    KA_LDX(ts%d, %s, %d, %d);", n, base.c_str(), ex_p(team.cut_id[n]), soff(ex_p(team.cut_id[n])));
      note("ts" + std::to_string(n));
      return;
    }
    if (leaf(n)) {
      fence_base();
      o.f("    KA_LDX(s%d, %s, %d, %d);", n, base.c_str(), n, soff(n));
      note("s" + std::to_string(n));
      if (loss_mode()) { o.f("    zx%d += ex2(fmaf(s%d, kLog2e, nm));", acc & 3, n); ++acc; }
      return;
    }
    const std::vector<int>& c = kids(n);
    for (int k : c) up(k);
    auto sn = [&](int k) { return std::string(top && is_cut(k) ? "ts" : "s") + std::to_string(k); };
    if (n != t.root) {
      std::string sum;
      for (size_t k = 0; k < c.size(); ++k) sum += (k ? " + " : "") + sn(c[k]);
      o.f("    const KF s%d = %s;", n, sum.c_str());
    }
    const int vid = plan.value_slot[n];
    if (c.size() == 1) return;  // a single child has probability 1: nothing to keep
    if (binary(n)) {
      char e[160];
      // t = m_a - m_b in log2 units
      snprintf(e, sizeof e, "fmaf(%s, %s, %s * %s)", sn(c[0]).c_str(), lit(kLog2e / t.cnt[c[0]]).c_str(), sn(c[1]).c_str(), lit(-kLog2e / t.cnt[c[1]]).c_str());
      if (v.mode == kSpecHardLoss) vset(vid, e);
      else vset(vid, std::string("signed_lo(") + e + ")");
      if (!spilled(vid)) note("nv" + std::to_string(vid));
      return;
    }
    // k-ary: softmax over the children (log2 domain)
    o.f("    {");
    for (int k : c) o.f("      const KF a%d = %s * %s;", k, sn(k).c_str(), lit(kLog2e / t.cnt[k]).c_str());
    std::string mx = "a" + std::to_string(c[0]);
    for (size_t k = 1; k < c.size(); ++k) mx = "fmaxf(" + mx + ", a" + std::to_string(c[k]) + ")";
    o.f("      const KF mx = %s;", mx.c_str());
    if (v.mode == kSpecHardLoss) {
      for (size_t k = 0; k < c.size(); ++k) o.f("      km%d = a%d - mx;", vid + (int)k, c[k]);
    } else {
      std::string sum;
      for (size_t k = 0; k < c.size(); ++k) {
        o.f("      const KF q%d = ex2(a%d - mx);", c[k], c[k]);
        sum += (k ? " + q" : "q") + std::to_string(c[k]);
      }
      o.f("      const KF r = rcp(%s);", sum.c_str());
      for (size_t k = 0; k < c.size(); ++k) o.f("      km%d = q%d * r;", vid + (int)k, c[k]);
    }
    o.f("    }");
    for (size_t k = 0; k < c.size(); ++k)
      if (spilled(vid + (int)k)) stx(spill_slot(vid + (int)k), "km" + std::to_string(vid + (int)k));
  }
  // k-ary values of the current part are declared before the blocks that set them
  void declare_kary() {
    for (int n = t.C; n < t.NV; ++n) {
      const std::vector<int>& c = kids(n);
      if (c.size() <= 2 || team.part_of_node[n] != part) continue;
      for (size_t k = 0; k < c.size(); ++k) o.f("    KF km%d;", plan.value_slot[n] + (int)k);
    }
  }

  // ---- pass 2 (soft): pre-order path probabilities ----
  // `pre` prefixes every variable name of the traversal (the loss makes two: "P" sums exp(P) and picks P[target]; "Q" is the
  // down half of the merged down/up gradient traversal).  what: 0 = soft forward (store P), 1 = sums only.
  void leaf_prob(int c, const char* pre, int what) {
    if (what == 0) {
      stx(c, std::string(pre) + std::to_string(c));
      note(std::string(pre) + std::to_string(c));
    } else {
      o.f("    { const KF e = %s(%s%d); z2%d += e; pry = sel_eq(posy, %d, %s%d, pry); }", (plan.knobs.expm & 1) ? "exp01m" : "exp01", pre, c, acc & 3, t.pos[c], pre, c);
      note("z2" + std::to_string(acc & 3));
      ++acc;
    }
  }
  // children's path probabilities of node n from the node's own (P: expression, "" = 1)
  void child_probs(int n, const std::string& P, const char* pre) {
    const std::vector<int>& c = kids(n);
    const int vid = plan.value_slot[n];
    if (c.size() == 1) {
      o.f("    const KF %s%d = %s;", pre, c[0], P.empty() ? "1.f" : P.c_str());
    } else if (binary(n)) {
      const std::string s = fenced(vget(vid));
      // the second traversal recomputes what the first one computed: without the opaque copy the compiler merges the two and
      // keeps every path probability of the first traversal alive (hundreds of spilled registers)
      if (pre[0] == 'Q') o.f("    KF %ss%d = %s; KA_OPAQUE(%ss%d);", pre, n, s.c_str(), pre, n);
      else o.f("    const KF %ss%d = %s;", pre, n, s.c_str());
      // s = -lo when the first child has the larger probability (1 - lo), +lo when it has the smaller one: with u = P or 0
      // accordingly, P*p(first) = P*s + u and P*p(second) = (P - u) - P*s -- two fused multiply-adds, each rounded once
      if (P.empty()) {
        o.f("    const KF %su%d = sel_neg(%ss%d, 1.f, 0.f);", pre, n, pre, n);
        o.f("    const KF %s%d = %ss%d + %su%d; const KF %s%d = (1.f - %su%d) - %ss%d;", pre, c[0], pre, n, pre, n, pre, c[1], pre, n, pre, n);
      } else {
        o.f("    const KF %su%d = sel_neg(%ss%d, %s, 0.f);", pre, n, pre, n, P.c_str());
        o.f("    const KF %s%d = fmaf(%s, %ss%d, %su%d); const KF %s%d = fmaf(-%s, %ss%d, %s - %su%d);", pre, c[0], P.c_str(), pre, n, pre, n,
            pre, c[1], P.c_str(), pre, n, P.c_str(), pre, n);
      }
    } else {
      for (size_t k = 0; k < c.size(); ++k) {
        const std::string pk = kget(vid + (int)k);
        if (pre[0] == 'Q') {
          o.f("    KF %sk%d = %s; KA_OPAQUE(%sk%d);", pre, c[k], pk.c_str(), pre, c[k]);
          if (P.empty()) o.f("    const KF %s%d = %sk%d;", pre, c[k], pre, c[k]);
          else o.f("    const KF %s%d = %s * %sk%d;", pre, c[k], P.c_str(), pre, c[k]);
        } else if (P.empty()) o.f("    const KF %s%d = %s;", pre, c[k], pk.c_str());
        else o.f("    const KF %s%d = %s * %s;", pre, c[k], P.c_str(), pk.c_str());
      }
    }
  }
  void down(int n, const std::string& P, const char* pre, int what) {
    child_probs(n, P, pre);
    for (int k : kids(n)) {
      if (top && is_cut(k)) { stx(ex_p(team.cut_id[k]), pre + std::to_string(k)); note(pre + std::to_string(k)); }
      else if (leaf(k)) leaf_prob(k, pre, what);
      else down(k, pre + std::to_string(k), pre, what);
    }
  }

  // ---- pass 3 (soft loss): one traversal, path probabilities on the way down, on the way up the subtree sums
  // H[v] = sum over the leaves under v of dL/dP[leaf] * P[leaf] (up to the factor cA) and per node the gradient of its child
  // logits dm[c] = H[c] - p[c]*H[v], which REPLACES the node's probability value (register or slot).  Returns the name of
  // H[n].  Nothing is stored per leaf: the row's x stays in its slots for the final pass.
  std::string leaf_g(int k, const std::string& Q) {
    o.f("    const KF G%d = %s * (%s(%s) - sel_eq(posy, %d, Z, 0.f));", k, Q.c_str(), (plan.knobs.expm & 2) ? "exp01m" : "exp01", Q.c_str(), t.pos[k]);
    return "G" + std::to_string(k);
  }
  void node_grad(int n, const std::string& H, const std::vector<std::string>& h, bool opaque) {
    const std::vector<int>& c = kids(n);
    const int vid = plan.value_slot[n];
    if (binary(n)) {
      std::string sp = opaque ? "Qs" + std::to_string(n) : vget(vid);
      // p(first) = s + (1 or 0), see child_probs: H[first] - p(first)*H = (H[first] - (H or 0)) - s*H
      o.f("    const KF pa%d = %s - sel_neg(%s, %s, 0.f);", n, h[0].c_str(), sp.c_str(), H.c_str());
      char e[128];
      snprintf(e, sizeof e, "cA * fmaf(-%s, %s, pa%d)", sp.c_str(), H.c_str(), n);
      dset(vid, e);
    } else {
      for (size_t k = 0; k < c.size(); ++k) {
        const std::string pk = opaque ? "Qk" + std::to_string(c[k]) : kget(vid + (int)k);
        char e[128];
        snprintf(e, sizeof e, "cA * fmaf(-%s, %s, %s)", pk.c_str(), H.c_str(), h[k].c_str());
        dset(vid + (int)k, e);
      }
    }
  }
  std::string updown(int n, const std::string& P) {
    const std::vector<int>& c = kids(n);
    child_probs(n, P, "Q");
    std::vector<std::string> h;
    for (int k : c) {
      if (leaf(k)) h.push_back(leaf_g(k, "Q" + std::to_string(k)));
      else h.push_back(updown(k, "Q" + std::to_string(k)));
    }
    const std::string H = "G" + std::to_string(n);
    if (c.size() == 1) {
      o.f("    const KF %s = %s;", H.c_str(), h[0].c_str());
      return H;
    }
    std::string sum;
    for (size_t k = 0; k < h.size(); ++k) sum += (k ? " + " : "") + h[k];
    o.f("    const KF %s = %s;", H.c_str(), sum.c_str());
    note(H);
    node_grad(n, H, h, true);
    return H;
  }
  // top tree, up half only: the sums of the cut subtrees come from their owners
  std::string top_h(int n) {
    const std::vector<int>& c = kids(n);
    std::vector<std::string> h;
    for (int k : c) {
      if (is_cut(k)) { fence_base(); h.push_back(ldx(ex_h(team.cut_id[k]))); }
      else h.push_back(top_h(k));
    }
    const std::string H = "G" + std::to_string(n);
    std::string sum;
    for (size_t k = 0; k < h.size(); ++k) sum += (k ? " + " : "") + h[k];
    o.f("    const KF %s = %s;", H.c_str(), sum.c_str());
    note(H);
    if (c.size() > 1) node_grad(n, H, h, false);
    return H;
  }
  // leaf c: dL/dx[c] = A[c] + cx * softmax(x)[c]  (the one-hot of the target is subtracted afterwards), written over x
  void leaf_out(int c, const std::string& A) {
    const std::string xv = ldx(c);
    char e[128];
    snprintf(e, sizeof e, "fmaf(cx, ex2(fmaf(%s, kLog2e, nl)), %s)", xv.c_str(), A.c_str());
    stx(c, e);
  }

  // ---- pass 4 (soft loss): pre-order, A[c] = A[v] + dm[c]/#leaves(c) ----
  void back(int n, const std::string& A) {
    const std::vector<int>& c = kids(n);
    const int vid = plan.value_slot[n];
    const char* ap = top ? "B" : "A";  // the top tree's names must not collide with warp 0's own subtrees
    auto an = [&](int k) { return std::string(ap) + std::to_string(k); };
    auto fma_a = [&](int child, const std::string& d) {
      const std::string inv = lit(1.0 / t.cnt[child]);
      if (A.empty()) o.f("    const KF %s%d = %s * %s;", ap, child, d.c_str(), inv.c_str());
      else o.f("    const KF %s%d = fmaf(%s, %s, %s);", ap, child, d.c_str(), inv.c_str(), A.c_str());
    };
    if (c.size() == 1) {
      o.f("    const KF %s%d = %s;", ap, c[0], A.empty() ? "0.f" : A.c_str());
    } else if (binary(n)) {
      const std::string d = fenced(dget(vid));
      fma_a(c[0], d);
      fma_a(c[1], "-" + d);
    } else {
      for (size_t k = 0; k < c.size(); ++k) fma_a(c[k], dget(vid + (int)k));
    }
    note(an(c.back()));
    for (int ck : c) {
      if (top && is_cut(ck)) stx(ex_h(team.cut_id[ck]), an(ck));
      else if (leaf(ck)) leaf_out(ck, an(ck));
      else back(ck, an(ck));
    }
  }

  // ---- hard loss: one pre-order pass (loss along the target's path + its gradient) ----
  // on: expression of the node's on-path indicator ("" = 1); A as above
  void hard(int n, const std::string& on, const std::string& A) {
    const std::vector<int>& c = kids(n);
    const int vid = plan.value_slot[n];
    const std::string onv = on.empty() ? "1.f" : on;
    const char* ap = top ? "B" : "A";    // the top tree's names must not collide with warp 0's own subtrees
    const char* op = top ? "ton" : "on";
    auto an = [&](int k) { return std::string(ap) + std::to_string(k); };
    auto onn = [&](int k) { return std::string(op) + std::to_string(k); };
    auto fma_a = [&](int child, const std::string& d) {
      const std::string inv = lit(1.0 / t.cnt[child]);
      if (A.empty()) o.f("    const KF %s%d = %s * %s;", ap, child, d.c_str(), inv.c_str());
      else o.f("    const KF %s%d = fmaf(%s, %s, %s);", ap, child, d.c_str(), inv.c_str(), A.c_str());
    };
    if (c.size() == 1) {
      o.f("    const KF %s%d = %s; const KF %s%d = %s;", op, c[0], onv.c_str(), ap, c[0], A.empty() ? "0.f" : A.c_str());
    } else if (binary(n)) {
      const int a = c[0], b = c[1];
      o.f("    const KF %s%d = sel_lt(posy, %d, %s, 0.f); const KF %s%d = %s - %s%d;", op, a, t.end[a], onv.c_str(), op, b, onv.c_str(), op, a);
      o.f("    const KF ht%d = %s; const KF hq%d = ex2(-fabsf(ht%d)); const KF hd%d = 1.f + hq%d; const KF hr%d = rcp(hd%d);",
          n, fenced(vget(vid)).c_str(), n, n, n, n, n, n);
      o.f("    const KF pa%d = sel_ge0(ht%d, hr%d, hq%d * hr%d);", n, n, n, n, n);
      // -log2 p(child) = log2(1 + 2^-|t|) + max(-t, 0) for the first child, + max(t, 0) for the second
      o.f("    lp = fmaf(%s, lg2(hd%d), lp); lp = fmaf(%s%d, fmaxf(-ht%d, 0.f), lp); lp = fmaf(%s%d, fmaxf(ht%d, 0.f), lp);",
          onv.c_str(), n, op, a, n, op, b, n);
      o.f("    const KF d%d = nkg * fmaf(-pa%d, %s, %s%d);", n, n, onv.c_str(), op, a);
      fma_a(a, "d" + std::to_string(n));
      fma_a(b, "-d" + std::to_string(n));
    } else {
      std::string sum;
      for (size_t k = 0; k < c.size(); ++k) {
        const int ck = c[k];
        o.f("    const KF %s%d = sel_in(posy, %d, %d, %s, 0.f);", op, ck, t.beg[ck], t.cnt[ck], onv.c_str());
        o.f("    const KF km_%d = %s; const KF kq_%d = ex2(km_%d);", ck, kget(vid + (int)k).c_str(), ck, ck);
        sum += (k ? " + kq_" : "kq_") + std::to_string(ck);
      }
      o.f("    const KF hs%d = %s; const KF hr%d = rcp(hs%d);", n, sum.c_str(), n, n);
      o.f("    lp = fmaf(%s, lg2(hs%d), lp);", onv.c_str(), n);
      for (int ck : c) {
        o.f("    lp = fmaf(-%s%d, km_%d, lp);", op, ck, ck);
        o.f("    const KF dk%d = nkg * fmaf(-kq_%d * hr%d, %s, %s%d);", ck, ck, n, onv.c_str(), op, ck);
        fma_a(ck, "dk" + std::to_string(ck));
      }
    }
    for (int ck : c) {
      note(an(ck));
      if (top && is_cut(ck)) {
        stx(ex_p(team.cut_id[ck]), onn(ck));
        stx(ex_h(team.cut_id[ck]), an(ck));
      } else if (leaf(ck)) {
        leaf_out(ck, an(ck));
      } else {
        hard(ck, onn(ck), an(ck));
      }
    }
  }

  void store_out(const char* ptr, int i, int vec, bool f32, const std::vector<std::string>& vals) {
    // vals: fp32 expressions of the elements of vector i
    std::vector<std::string> words;
    if (f32) {
      for (const std::string& e : vals) words.push_back("f2u(" + e + ")");
    } else if (vec == 2) {
      words.push_back(std::string(v.dtype == NBDT_BF16 ? "pack_bf16(" : "pack_f16(") + vals[0] + ", 0.f)");
    } else {
      for (size_t k = 0; k + 1 < vals.size(); k += 2)
        words.push_back(std::string(v.dtype == NBDT_BF16 ? "pack_bf16(" : "pack_f16(") + vals[k] + ", " + vals[k + 1] + ")");
    }
    std::string args;
    for (size_t k = 0; k < words.size(); ++k) args += ", " + words[k];
    o.f("    st%d(%s + %d%s);", vec, ptr, i * vec, args.c_str());
  }

  // one vector of the row block in the generic (strided / partial block) sweep
  void coop_load(const char* name) {
    switch (v.vec) {
      case 16: o.f("        const uint4 %s0 = ld16(gp);", name); break;
      case 8: o.f("        const uint2 %s0 = ld8(gp);", name); break;
      case 4: o.f("        const u32 %s0 = ld4(gp);", name); break;
      default: o.f("        const u32 %s0 = ld2(gp);", name); break;
    }
  }
  // advance (row, col) of this thread's vector by one sweep iteration (KA_TPB vectors of pv elements)
  void coop_step(int pv) {
    o.f("      col += (KA_TPB * %d) %% KA_C; row += (KA_TPB * %d) / KA_C;", pv, pv);
    o.f("      if (col >= KA_C) { col -= KA_C; ++row; }");
  }

  // ---- the straight-line code of one warp of the team for one row ----
  void emit_part(int w) {
    const int C = t.C, W = team.W;
    part = w; top = false; acc = 0; since = 0; lock_since = 0; last.clear(); base = "xsb";
    const std::vector<int>& cuts = team.cuts_of_part[w];
    std::vector<int> own;  // own leaves, DFS order
    for (int k : cuts)
      for (int pos = t.beg[team.cut_nodes[k]]; pos < t.end[team.cut_nodes[k]]; ++pos) own.push_back(t.leaf_at[pos]);
    const bool first = w == 0;
    o.f("// warp %d of %d: %d cut subtrees, %d leaves%s", w, W, (int)cuts.size(), (int)own.size(), first && teamed() ? " + the top tree" : "");
    o.f("KA_DEV void ka_part%d(const KAParams& p, KY yy, KB live, KA_XSB xsb, u32 zr, float& loss_acc) {", w);
    if (loss_mode()) {
      // rows with an out-of-range target are ignored: zero loss, zero gradient (nn.CrossEntropyLoss's ignore_index rows; the
      // reference's class_index_to_child_index has no entry for them either).  They run the arithmetic on target 0 and are
      // masked at the two places where a row has an effect.
      o.f("    const KB valid = ka_valid(yy);");
      o.f("    const KI y = ka_target(yy, valid);");
      o.f("    const KI posy = ka_lookup(ka_pos, y);");
      o.f("    const float kx = p.w_xent * p.inv_rows, kt = p.w_tree * p.inv_rows;");
      if (first) o.f("    KA_LDX_DYN(xy, y);");
      // max over the own leaves
      o.f("    KF m0 = -3.0e38f, m1 = m0, m2 = m0, m3 = m0;");
      for (size_t i = 0; i < own.size(); ++i) {
        const std::string xv = ldx(own[i]);
        o.f("    m%d = fmaxf(m%d, %s);", (int)(i & 3), (int)(i & 3), xv.c_str());
      }
      o.f("    const KF mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));");
      o.f("    const KF nm = -mx * kLog2e;");
      o.f("    KF zx0 = 0.f, zx1 = 0.f, zx2 = 0.f, zx3 = 0.f;");
    }
    // ---- pass 1 over the own subtrees ----
    declare_kary();
    for (int k : cuts) {
      const int r = team.cut_nodes[k];
      up(r);
      if (teamed()) stx(ex_p(k), "s" + std::to_string(r));
    }
    if (loss_mode()) {
      o.f("    const KF zxw = (zx0 + zx1) + (zx2 + zx3);");
      if (teamed()) { stx(ex_w(0, w), "mx"); stx(ex_w(1, w), "zxw"); }
    }
    o.f("    KA_BAR();");
    // ---- the whole row's log-sum-exp from the warps' (max, sum) pairs ----
    if (loss_mode()) {
      if (teamed()) {
        std::vector<std::string> ms, zs;
        for (int u = 0; u < W; ++u) { ms.push_back(ldx(ex_w(0, u))); zs.push_back(ldx(ex_w(1, u))); }
        std::string mg = ms[0];
        for (int u = 1; u < W; ++u) mg = "fmaxf(" + mg + ", " + ms[u] + ")";
        o.f("    const KF mxg = %s;", mg.c_str());
        o.f("    KF zxg = 0.f;");
        for (int u = 0; u < W; ++u) o.f("    zxg = fmaf(%s, ex2((%s - mxg) * kLog2e), zxg);", zs[u].c_str(), ms[u].c_str());
      } else {
        o.f("    const KF mxg = mx, zxg = zxw;");
      }
      o.f("    const KF l2zx = lg2(zxg);");
      o.f("    const KF lse1 = fmaf(l2zx, kLn2, mxg);");
      o.f("    const float cx = kx * p.grad_scale;");
      o.f("    const KF nl = -fmaf(mxg, kLog2e, l2zx);");
    }
    // ---- top tree, pass 1 and the way down (warp 0) ----
    if (teamed()) {
      if (first) {
        top = true;
        up(t.root);
        if (v.mode == kSpecHardLoss) {
          o.f("    KF lp = 0.f;");
          o.f("    const float nkg = -kt * p.grad_scale;");
          hard(t.root, "", "");
        } else {
          down(t.root, "", "T", 0);
        }
        top = false;
      }
      o.f("    KA_BAR();");
    }
    // ---- own subtrees ----
    if (v.mode == kSpecSoftFwd) {
      for (int k : cuts) {
        const int r = team.cut_nodes[k];
        std::string P = teamed() ? ldx(ex_p(k)) : "";
        if (teamed()) { o.f("    const KF P%d = %s;", r, P.c_str()); P = "P" + std::to_string(r); }
        if (leaf(r)) { if (!teamed()) o.f("    const KF P%d = 1.f;", r); leaf_prob(r, "P", 0); }
        else down(r, P, "P", 0);
      }
      o.f("}");
      return;
    }
    if (v.mode == kSpecHardLoss) {
      if (!teamed()) {
        o.f("    KF lp = 0.f;");
        o.f("    const float nkg = -kt * p.grad_scale;");
        hard(t.root, "", "");
      } else {
        if (!first) { o.f("    KF lp = 0.f;"); o.f("    const float nkg = -kt * p.grad_scale;"); }
        for (int k : cuts) {
          const int r = team.cut_nodes[k];
          const std::string on = ldx(ex_p(k)), A = ldx(ex_h(k));
          o.f("    const KF on%d = %s; const KF A%d = %s;", r, on.c_str(), r, A.c_str());
          if (leaf(r)) leaf_out(r, "A" + std::to_string(r));
          else hard(r, "on" + std::to_string(r), "A" + std::to_string(r));
        }
        stx(ex_w(3, w), "lp");
        o.f("    KA_BAR();");
        if (first) {
          std::string s = "0.f";
          for (int u = 0; u < W; ++u) s += " + " + ldx(ex_w(3, u));
          o.f("    const KF lpg = %s;", s.c_str());
        }
      }
      if (first) o.f("    loss_acc += ka_hsum(sel(both(live, valid), kx * (lse1 - xy) + (kt * kLn2) * %s, 0.f));", teamed() ? "lpg" : "lp");
    } else {
      // soft loss: sum exp(P) and P[target] over the own leaves
      o.f("    KF z20 = 0.f, z21 = 0.f, z22 = 0.f, z23 = 0.f, pry = 0.f;");
      acc = 0;
      std::vector<std::string> Pr(cuts.size());
      for (size_t i = 0; i < cuts.size(); ++i) {
        const int k = cuts[i], r = team.cut_nodes[k];
        if (teamed()) { Pr[i] = ldx(ex_p(k)); o.f("    const KF P%d = %s;", r, Pr[i].c_str()); Pr[i] = "P" + std::to_string(r); }
        if (leaf(r)) { if (!teamed()) o.f("    const KF P%d = 1.f;", r); leaf_prob(r, "P", 1); }
        else down(r, Pr[i], "P", 1);
      }
      o.f("    const KF Zw = (z20 + z21) + (z22 + z23);");
      if (teamed()) {
        stx(ex_w(2, w), "Zw"); stx(ex_w(3, w), "pry");
        o.f("    KA_BAR();");
        std::string sz = "0.f", sp = "0.f";
        for (int u = 0; u < W; ++u) { sz += " + " + ldx(ex_w(2, u)); sp += " + " + ldx(ex_w(3, u)); }
        o.f("    const KF Z = %s; const KF pryg = %s;", sz.c_str(), sp.c_str());
      } else {
        o.f("    const KF Z = Zw, pryg = pry;");
      }
      if (first) o.f("    loss_acc += ka_hsum(sel(both(live, valid), kx * (lse1 - xy) + kt * (kLn2 * lg2(Z) - pryg), 0.f));");
      o.f("    const KF cA = kt * p.grad_scale * rcp(Z);");
      // the merged down/up traversal of the own subtrees
      for (size_t i = 0; i < cuts.size(); ++i) {
        const int k = cuts[i], r = team.cut_nodes[k];
        std::string Q = teamed() ? "Q" + std::to_string(r) : "";
        if (teamed()) o.f("    KF Q%d = %s; KA_OPAQUE(Q%d);", r, Pr[i].c_str(), r);
        std::string H;
        if (leaf(r)) { if (!teamed()) o.f("    const KF Q%d = 1.f;", r); H = leaf_g(r, "Q" + std::to_string(r)); }
        else H = updown(r, Q);
        if (teamed()) stx(ex_h(k), H);
      }
      if (teamed()) {
        o.f("    KA_BAR();");
        if (first) {  // top tree: sums up, gradient down to the cut roots
          top = true;
          top_h(t.root);
          back(t.root, "");
          top = false;
        }
        o.f("    KA_BAR();");
      }
      for (int k : cuts) {
        const int r = team.cut_nodes[k];
        std::string A = "";
        if (teamed()) { A = ldx(ex_h(k)); o.f("    const KF A%d = %s;", r, A.c_str()); A = "A" + std::to_string(r); }
        if (leaf(r)) leaf_out(r, A.empty() ? "0.f" : A);
        else back(r, A);
      }
    }
    // the slots of the own leaves now hold dL/dx except for the one-hot of the plain cross-entropy term
    if (teamed()) o.f("    KA_ADDX_DYN(ka_owned(ka_owner, y, %d), y, -cx);", w);
    else o.f("    KA_ADDX_DYN(ka_true(), y, -cx);");
    // an ignored row gets a zero gradient row (rare: a loop, not straight-line code)
    o.f("    if (!ka_all(valid)) {");
    if (teamed()) o.f("      for (int c = 0; c < KA_C; ++c) if (ka_owner[c] == %d) KA_ZERO_IF(ka_not(valid), c);", w);
    else o.f("      for (int c = 0; c < KA_C; ++c) KA_ZERO_IF(ka_not(valid), c);");
    o.f("    }");
    o.f("}");
  }

  // Class-major tensors (layout 1): lane = pixel.  Class c of the block's 32 pixels is 32 consecutive elements, so one warp-wide
  // load / store per class is one coalesced line, and every thread fills (reads) its own slot c.  The warps of a team share the
  // classes round-robin.  A block never straddles two images (hw % 32 == 0).
  void emit_sweeps_class_major() {
    const int W = team.W;
    const char* ldf = v.dtype == NBDT_F32 ? "u2f(ld4(src))" : (v.dtype == NBDT_BF16 ? "u2f(ld2(src) << 16)" : "h2f(ld2(src))");
    o.f("KA_DEV void ka_sweep_a(const KAParams& p, long long row0, int nrows, KA_XSB xsw, int tid) {");
    o.f("#ifdef KA_HOST\n    const int lane = 0, c_first = 0, c_step = 1;\n    const KA_XSB xsb = xsw;");
    o.f("#else\n    const int lane = tid & 31, c_first = tid >> 5, c_step = %d;\n    const KA_XSB xsb = xsw + lane * 4u;\n#endif", W);
    o.f("    const long long img = row0 / p.hw, pix = row0 - img * p.hw + lane;");
    o.f("    const char* xb = reinterpret_cast<const char*>(p.x) + (img * KA_C * p.hw + pix) * %d;", esz());
    o.f("    if (lane < nrows) {");
    o.f("#pragma unroll 8");
    o.f("      for (int c = c_first; c < KA_C; c += c_step) {");
    o.f("        const char* src = xb + (long long)c * p.hw * %d;", esz());
    o.f("        KA_STX_DYN(c, %s);", ldf);
    o.f("      }");
    o.f("    }");
    o.f("}");
    const bool fwd = v.mode == kSpecSoftFwd;
    const int oesz = fwd ? 4 : esz();
    o.f("KA_DEV void ka_sweep_b(const KAParams& p, long long row0, int nrows, KA_XSB xsw, int tid) {");
    o.f("#ifdef KA_HOST\n    const int lane = 0, c_first = 0, c_step = 1;\n    const KA_XSB xsb = xsw;");
    o.f("#else\n    const int lane = tid & 31, c_first = tid >> 5, c_step = %d;\n    const KA_XSB xsb = xsw + lane * 4u;\n#endif", W);
    o.f("    const long long img = row0 / p.hw, pix = row0 - img * p.hw + lane;");
    o.f("    char* ob = reinterpret_cast<char*>(%s) + (img * KA_C * p.hw + pix) * %d;", fwd ? "p.class_probs" : "p.grad", oesz);
    o.f("    if (lane < nrows) {");
    o.f("#pragma unroll 8");
    o.f("      for (int c = c_first; c < KA_C; c += c_step) {");
    o.f("        KA_LDX_DYN(q, c);");
    o.f("        char* dst = ob + (long long)c * p.hw * %d;", oesz);
    if (fwd || v.dtype == NBDT_F32) o.f("        st4(dst, f2u(q));");
    else o.f("        st2(dst, %s(q, 0.f));", v.dtype == NBDT_BF16 ? "pack_bf16" : "pack_f16");
    o.f("      }");
    o.f("    }");
    o.f("}");
  }

  std::string run() {
    const int C = t.C, W = team.W;
    const int pv = per_vec();
    o.s.reserve((size_t)C * 2600);
    U = v.mode == kSpecSoftFwd ? std::min(pv, v.ovec / 4) : pv;
    K = ((C / U) & 31) | 1;
    if (v.layout == 1) { U = 1 << 20; K = 1; }  // thread-private slots only: plain [class][lane]
    const int RB = 32 * R();  // rows per block
    const int T = plan.teams;
    o.f("#define KA_C %d\n#define KA_W %d\n#define KA_R %d\n#ifdef KA_HOST\n#define KA_T 1\n#else\n#define KA_T %d\n#endif", C, W, R(), T);
    o.f("#ifdef KA_HOST\n#define KA_TPB 1\n#else\n#define KA_TPB %d\n#endif\n#define KA_U %du\n#define KA_K %du", kTPB * W, U, K);
    o.s += kPrelude;
    auto table = [&](const char* type, const char* name, const std::vector<int>& vals) {
      std::string tab;
      for (size_t c = 0; c < vals.size(); ++c) tab += std::to_string(vals[c]) + (c + 1 < vals.size() ? "," : "");
      o.f("#ifdef KA_HOST\nstatic const %s %s[%d] = {%s};\n#else\n__device__ const %s %s[%d] = {%s};\n#endif", type, name, (int)vals.size(),
          tab.c_str(), type, name, (int)vals.size(), tab.c_str());
    };
    if (loss_mode()) {
      table("unsigned short", "ka_pos", t.pos);
      if (teamed()) {
        std::vector<int> owner(C);
        for (int c = 0; c < C; ++c) owner[c] = team.part_of_node[c];
        table("unsigned char", "ka_owner", owner);
      }
    }
    o.f("// hierarchy %016llx: %d leaves, %d inner nodes; mode %d dtype %d vec %d; %d node values; %d warps per row block, %d cut subtrees, %d top nodes",
        t.hash, C, t.N, v.mode, v.dtype, v.vec, plan.nvalues, W, team.ncut(), team.ntop);

    if (v.layout == 1) emit_sweeps_class_major();
    else {
    // ---- sweep A: the block's rows, coalesced, into the transposed slots ----
    o.f("KA_DEV void ka_sweep_a(const KAParams& p, long long row0, int nrows, KA_XSB xsw, int tid) {");
    o.f("    const char* xb = reinterpret_cast<const char*>(p.x) + row0 * p.ld * %d;", esz());
    o.f("    int row = (tid * %d) / KA_C, col = (tid * %d) %% KA_C;", pv, pv);
    const int nvecs = RB * C / pv;                          // vectors per block
    const int nit = (nvecs + kTPB * W - 1) / (kTPB * W);    // per thread
    // full block of contiguous rows (the common case): every load has an immediate offset from one pointer and a whole group
    // of loads is in flight before the first store needs its data
    o.f("#ifndef KA_HOST");
    o.f("    if (nrows == %d && p.ld == KA_C) {", RB);
    o.f("      const char* gp = xb + tid * %d;", v.vec);
    const int maxld = std::max(1, plan.knobs.maxld * 16 / v.vec);
    for (int g0 = 0; g0 < nit; g0 += maxld) {
      const int g1 = std::min(nit, g0 + maxld);
      for (int it = g0; it < g1; ++it) {
        const bool guard = (it + 1) * kTPB * W > nvecs;  // the last iteration may run past the block
        const char* ty = v.vec == 16 ? "uint4" : (v.vec == 8 ? "uint2" : "u32");
        const char* zero = v.vec == 16 ? "make_uint4(0u, 0u, 0u, 0u)" : (v.vec == 8 ? "make_uint2(0u, 0u)" : "0u");
        if (guard) o.f("      const %s lv%d = (tid + %d < %d) ? ld%d(gp + %d) : %s;", ty, it, it * kTPB * W, nvecs, v.vec, it * kTPB * W * v.vec, zero);
        else o.f("      const %s lv%d = ld%d(gp + %d);", ty, it, v.vec, it * kTPB * W * v.vec);
      }
      for (int it = g0; it < g1; ++it) {
        const bool guard = (it + 1) * kTPB * W > nvecs;
        o.f("      %s{ const KA_XSB a = KA_SLOT_ADDR(xsw, col, row);", guard ? ("if (row < " + std::to_string(RB) + ") ").c_str() : "");
        for (int j = 0; j < pv; ++j) o.f("        KA_STD(a, %d, %d, %s);", j, eoff(j), elem("lv", it * pv + j).c_str());
        o.f("      }");
        coop_step(pv);
      }
    }
    o.f("      return;");
    o.f("    }");
    o.f("#endif");
    if (plan.knobs.unroll > 0) o.f("#pragma unroll %d", plan.knobs.unroll); else o.f("#pragma unroll");
    o.f("#ifdef KA_HOST\n    const int nit = %d;\n#else\n    const int nit = %d;\n#endif", R() * C / pv, nit);
    o.f("    for (int it = 0; it < nit; ++it) {");
    o.f("      if (row < nrows) {");
    o.f("        const char* gp = xb + ((long long)row * p.ld + col) * %d;", esz());
    coop_load("lv");
    o.f("        const KA_XSB a = KA_SLOT_ADDR(xsw, col, row);");
    for (int j = 0; j < pv; ++j) o.f("        KA_STD(a, %d, %d, %s);", j, eoff(j), elem("lv", j).c_str());
    o.f("      }");
    coop_step(pv);
    o.f("    }");
    o.f("}");
    }

    // ---- the row code, one function per warp of the team ----
    for (int w = 0; w < W; ++w) emit_part(w);
    o.f("#ifdef KA_HOST");
    o.f("static void ka_part(int w, const KAParams& p, KY yy, KB live, KA_XSB xsb, u32 zr, float& loss_acc) {");
    o.f("  switch (w) {");
    for (int w = 0; w < W; ++w) o.f("    case %d: ka_part%d(p, yy, live, xsb, zr, loss_acc); break;", w, w);
    o.f("  }");
    o.f("}");
    o.f("#endif");

    // ---- sweep B: coalesced write of the block's gradient rows / class probabilities (the slots hold the final values) ----
    if (v.layout == 0) {
    o.f("KA_DEV void ka_sweep_b(const KAParams& p, long long row0, int nrows, KA_XSB xsw, int tid) {");
    {
      const bool fwd = v.mode == kSpecSoftFwd;
      const int ov = fwd ? v.ovec : v.vec;          // bytes per output vector
      const int po = fwd ? v.ovec / 4 : pv;         // elements per output vector
      const int oesz = fwd ? 4 : esz();
      const int nvo = RB * C / po;
      const int nito = (nvo + kTPB * W - 1) / (kTPB * W);
      if (fwd) o.f("    char* ob = reinterpret_cast<char*>(p.class_probs + row0 * KA_C); const long long old = KA_C;");
      else o.f("    char* ob = reinterpret_cast<char*>(p.grad) + row0 * p.ld * %d; const long long old = p.ld;", oesz);
      o.f("    int row = (tid * %d) / KA_C, col = (tid * %d) %% KA_C;", po, po);
      o.f("#ifndef KA_HOST");
      o.f("    if (nrows == %d && old == KA_C) {", RB);
      o.f("      char* op = ob + tid * %d;", ov);
      for (int it = 0; it < nito; ++it) {
        const bool guard = (it + 1) * kTPB * W > nvo;
        o.f("      %s{ const KA_XSB a = KA_SLOT_ADDR(xsw, col, row);", guard ? ("if (row < " + std::to_string(RB) + ") ").c_str() : "");
        std::vector<std::string> vals;
        for (int j = 0; j < po; ++j) {
          o.f("        KA_LDD(q%d, a, %d, %d);", j, j, eoff(j));
          vals.push_back("q" + std::to_string(j));
        }
        store_out("    op", it * kTPB * W, ov, fwd || v.dtype == NBDT_F32, vals);
        o.f("      }");
        coop_step(po);
      }
      o.f("      return;");
      o.f("    }");
      o.f("#endif");
      if (plan.knobs.unroll > 0) o.f("#pragma unroll %d", plan.knobs.unroll); else o.f("#pragma unroll");
      o.f("#ifdef KA_HOST\n    const int nit = %d;\n#else\n    const int nit = %d;\n#endif", R() * C / po, nito);
      o.f("    for (int it = 0; it < nit; ++it) {");
      o.f("      if (row < nrows) {");
      o.f("        const KA_XSB a = KA_SLOT_ADDR(xsw, col, row);");
      std::vector<std::string> vals;
      for (int j = 0; j < po; ++j) {
        o.f("        KA_LDD(q%d, a, %d, %d);", j, j, eoff(j));
        vals.push_back("q" + std::to_string(j));
      }
      o.f("        char* op = ob + ((long long)row * old + col) * %d;", oesz);
      store_out("    op", 0, ov, fwd || v.dtype == NBDT_F32, vals);
      o.f("      }");
      coop_step(po);
      o.f("    }");
    }
    o.f("}");
    }

    // ---- kernel: one team per CTA, 32 rows per iteration ----
    o.f("#ifndef KA_HOST");
    // CTA = KA_T teams of KA_W warps; `tid` is the thread's index within its team, which owns its own shared-memory block and its own
    // row block; all teams make the same number of trips (a team past the end of the batch runs on zero live rows) because the
    // warps that share a part meet at the lock-step barriers
    const int team_bytes = plan.smem_bytes / T;
    o.f("extern \"C\" __global__ void __launch_bounds__(KA_TPB * KA_T, %d) nbdt_tree_ka(const KAParams p) {", min_blocks());
    o.f("  extern __shared__ float ka_smem[];");
    o.f("  const int team = threadIdx.x / KA_TPB, tid = threadIdx.x - team * KA_TPB, lane = tid & 31;");
    o.f("  const u32 xsw = (u32)__cvta_generic_to_shared(ka_smem) + team * %du;", team_bytes);
    o.f("#if KA_T > 1");
    o.f("#define KA_TEAM_SYNC() asm volatile(\"bar.sync %%0, %%1;\" :: \"r\"(1u + team), \"n\"(KA_TPB) : \"memory\")");
    o.f("#elif KA_W > 1");
    o.f("#define KA_TEAM_SYNC() __syncthreads()");
    o.f("#else");
    o.f("#define KA_TEAM_SYNC() __syncwarp()");
    o.f("#endif");
    o.f("  const u32 xsb = xsw + ((lane * KA_K) & 31u) * %du;", 4 * R());
    o.f("  float loss_acc = 0.f;");
    o.f("  const u32 zr = (u32)((unsigned long long)p.B >> 48);  // zero at run time, unknown at compile time (scheduling fences)");
    o.f("  for (long long blk0 = (long long)blockIdx.x * KA_T * %d; blk0 < p.B; blk0 += (long long)gridDim.x * KA_T * %d) {", RB, RB);
    o.f("    const long long row0 = blk0 + team * %d;", RB);
    o.f("    const int nrows = (int)(p.B - row0 < %d ? (p.B - row0 > 0 ? p.B - row0 : 0) : %d);", RB, RB);
    if (R() == 1) {
      o.f("    KY yy = 0; const KB live = lane < nrows;");
      if (loss_mode()) o.f("    if (live) yy = p.targets[row0 + lane];");
    } else {
      o.f("    KY yy; yy.a = 0; yy.b = 0; const KB live = mkb(2 * lane < nrows, 2 * lane + 1 < nrows);");
      if (loss_mode()) o.f("    if (live.a) yy.a = p.targets[row0 + 2 * lane]; if (live.b) yy.b = p.targets[row0 + 2 * lane + 1];");
    }
    if (plan.knobs.prefetch > 0) {
      // the lines of a block further ahead go to L2 now, so that its sweep A pays an L2 hit instead of an HBM access
      o.f("    {");
      o.f("      const long long pr0 = row0 + %dll * gridDim.x * KA_T * %d;", plan.knobs.prefetch, RB);
      o.f("      if (pr0 < p.B && p.ld <= 2 * KA_C) {");
      o.f("        const long long pn = p.B - pr0 < %d ? p.B - pr0 : %d;", RB, RB);
      o.f("        const char* pb = reinterpret_cast<const char*>(p.x) + pr0 * p.ld * %d;", esz());
      o.f("        const long long plen = ((pn - 1) * p.ld + KA_C) * %d;", esz());
      o.f("        for (long long po = tid * 128; po < plen; po += KA_TPB * 128) asm volatile(\"prefetch.global.L2 [%%0];\" :: \"l\"(pb + po));");
      o.f("      }");
      o.f("    }");
    }
    o.f("    if (nrows > 0) ka_sweep_a(p, row0, nrows, xsw, tid);");
    o.f("    KA_TEAM_SYNC();");
    if (W > 1 || T > 1) {
      o.f("    switch (tid >> 5) {");
      for (int w = 0; w < W; ++w) o.f("      case %d: ka_part%d(p, yy, live, xsb, zr, loss_acc); break;", w, w);
      o.f("    }");
    } else {
      // a single warp per CTA has no barriers inside the row function: threads without a row skip it
      o.f("    if (lane * %d < nrows) ka_part0(p, yy, live, xsb, zr, loss_acc);", R());
    }
    o.f("    KA_TEAM_SYNC();");
    if (loss_mode()) o.f("    if (p.grad && nrows > 0) ka_sweep_b(p, row0, nrows, xsw, tid);");
    else o.f("    if (nrows > 0) ka_sweep_b(p, row0, nrows, xsw, tid);");
    o.f("    KA_TEAM_SYNC();");
    o.f("  }");
    if (loss_mode()) {
      o.f("  if (tid < 32) {");
      o.f("    for (int s = 16; s > 0; s >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, s);");
      o.f("    if (tid == 0) atomicAdd(p.loss, loss_acc);");
      o.f("  }");
    }
    o.f("}");
    o.f("#endif");
    return o.s;
  }
};

// cut the hierarchy for a team of W warps
Team make_team(const KTree& t, int W) {
  Team tm;
  tm.W = std::max(1, W);
  tm.cut_id.assign(t.NV, -1);
  tm.part_of_node.assign(t.NV, 0);
  tm.is_top.assign(t.NV, 0);
  if (tm.W == 1) {
    tm.cut_id[t.root] = 0;
    tm.cut_nodes.push_back(t.root);
    tm.part_of_cut.push_back(0);
    tm.cuts_of_part.assign(1, std::vector<int>(1, 0));
    return tm;
  }
  const int limit = std::max(1, (t.C + 2 * tm.W - 1) / (2 * tm.W));  // leaves per cut subtree at most
  std::vector<int> st(1, t.root);
  while (!st.empty()) {  // DFS, children pushed in reverse so that cut subtrees come out in DFS order
    const int n = st.back();
    st.pop_back();
    if (t.cnt[n] <= limit || n < t.C) {
      tm.cut_id[n] = (int)tm.cut_nodes.size();
      tm.cut_nodes.push_back(n);
      continue;
    }
    tm.is_top[n] = 1;
    ++tm.ntop;
    const std::vector<int>& c = t.ch[n - t.C];
    for (size_t k = c.size(); k-- > 0;) st.push_back(c[k]);
  }
  // longest-processing-time packing; warp 0 starts with the top tree on its account
  std::vector<int> order(tm.cut_nodes.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return t.cnt[tm.cut_nodes[a]] > t.cnt[tm.cut_nodes[b]]; });
  std::vector<int> load(tm.W, 0);
  load[0] = tm.ntop;
  tm.part_of_cut.assign(tm.cut_nodes.size(), 0);
  for (int k : order) {
    int best = 0;
    for (int w = 1; w < tm.W; ++w) if (load[w] < load[best]) best = w;
    tm.part_of_cut[k] = best;
    load[best] += t.cnt[tm.cut_nodes[k]];
  }
  tm.cuts_of_part.assign(tm.W, std::vector<int>());
  for (size_t k = 0; k < tm.cut_nodes.size(); ++k) tm.cuts_of_part[tm.part_of_cut[k]].push_back((int)k);
  // owner of every node
  for (size_t k = 0; k < tm.cut_nodes.size(); ++k) {
    std::vector<int> q(1, tm.cut_nodes[k]);
    while (!q.empty()) {
      const int n = q.back();
      q.pop_back();
      tm.part_of_node[n] = tm.part_of_cut[k];
      if (n >= t.C) for (int c : t.ch[n - t.C]) q.push_back(c);
    }
  }
  return tm;
}

int auto_rows(const KTree& t, const Knobs& knobs) {
  if (knobs.rows > 0) return knobs.rows;
  (void)t;
  return 1;
}

int auto_warps(const KTree& t) {
  // measured on B200 (gpurun_out/tune_v5.jsonl): one warp per row block is fastest up to ~128 leaves (12 resident warps
  // per SM); at 200 leaves one warp needs 255 registers and 30 KB (7 warps per SM) and a team of 4 wins
  if (t.C <= 128) return 1;
  if (t.C <= 256) return 4;
  int w = 4;
  while (w < 16 && t.C > 64 * w) w *= 2;
  return w;
}

Plan make_plan(const KTree& t, int max_smem, const Knobs& knobs) {
  Plan p;
  p.knobs = knobs;
  p.default_knobs = knobs.is_default();
  p.team = make_team(t, knobs.warps > 0 ? knobs.warps : auto_warps(t));
  const int W = p.team.W;
  p.threads = kTPB * W;
  p.value_slot.assign(t.NV, -1);
  // values in post-order (the order pass 1 produces them)
  std::vector<int> order;  // inner nodes, post-order
  {
    struct Frame { int v; size_t k; };
    std::vector<Frame> st;
    st.push_back({t.root, 0});
    while (!st.empty()) {
      Frame& f = st.back();
      const std::vector<int>& c = t.ch[f.v - t.C];
      if (f.k == c.size()) { order.push_back(f.v); st.pop_back(); continue; }
      const int u = c[f.k++];
      if (u >= t.C) st.push_back({u, 0});
    }
  }
  if (p.knobs.expm < 0) p.knobs.expm = W == 1 ? 1 : 3;
  // Scheduling fences trade registers for instruction-level parallelism.  One warp per block, 65..128 leaves: no fences and at
  // most 8 resident warps (255 registers, a few dozen bytes of local spills) beat fences every 8 nodes at 168 registers and 12
  // warps (100 leaves: 42.8 % vs 39.7 % of the HBM peak); teams want the occupancy and a window of 16 nodes (200 leaves: 24.5 %
  // vs 23.5 %); small hierarchies keep the round-1 setting.
  const bool wide = W == 1 && t.C > 64;
  p.wide = wide && p.knobs.fence < 0 && knobs.minb == 0;
  if (p.knobs.fence < 0) p.knobs.fence = wide ? 0 : (W > 1 && t.C <= 256 ? 16 : 8);
  p.rows = auto_rows(t, knobs);
  const int R = p.rows;
  // register budget per thread: what a full SM of resident teams allows, minus the working set of the passes; a value of a
  // two-row thread is a register pair
  const int regcap = (W == 1 ? knobs.regcap : std::min(knobs.regcap, std::max(16, 65536 / (kTPB * W) - 72))) / R;
  std::vector<int> in_regs(W, 0), nspill(W, 0);
  for (int n : order) {
    const size_t k = t.ch[n - t.C].size();
    if (k <= 1) continue;
    p.value_slot[n] = p.nvalues;
    const int cnt = k == 2 ? 1 : (int)k;
    const int w = p.team.part_of_node[n];
    for (int j = 0; j < cnt; ++j) {
      p.value_part.push_back(w);
      if (in_regs[w] < regcap) { ++in_regs[w]; p.value_spill.push_back(-1); }
      else p.value_spill.push_back(nspill[w]++);
    }
    p.nvalues += cnt;
  }
  int maxregs = 0;
  for (int w = 0; w < W; ++w) { p.maxspill = std::max(p.maxspill, nspill[w]); maxregs = std::max(maxregs, in_regs[w]); }
  const int slots = t.C + W * p.maxspill + 2 * p.team.ncut() + 4 * W;
  p.smem_bytes = (t.C * 32 + t.C + 36 + (slots - t.C) * 32) * 4 * R;
  const int regs = std::min(255, (maxregs * R + (R == 1 ? 64 : 96) + 7) / 8 * 8);
  int by_regs = 65536 / (kTPB * W * regs);
  int by_smem = (max_smem > 0 ? max_smem : 227 * 1024) / std::max(1, p.smem_bytes + 1024);
  p.min_blocks = std::max(1, std::min(32, std::min(by_regs, by_smem)));
  if (knobs.minb > 0) p.min_blocks = knobs.minb;
  // several teams per CTA: as many as would have been resident as separate CTAs, barrier ids 1..teams (team) and
  // teams+1..teams+W (lock step) must stay below 16
  p.teams = 1;
  if (knobs.teams > 1) {
    const int resident = p.wide ? std::min(p.min_blocks, 8) : p.min_blocks;  // CTAs per SM the single-team kernel would have had
    int T = std::min(knobs.teams, std::max(1, resident));
    while (T > 1 && (1 + T + W > 16 || (long long)T * p.smem_bytes > (max_smem > 0 ? max_smem : 227 * 1024) - 1024)) --T;
    p.teams = T;
    if (T > 1) {
      p.min_blocks = std::max(1, resident / T);
      p.smem_bytes *= T;
      p.threads *= T;
    }
  }
  return p;
}

// ------------------------------------------------------------------------------------------------------------------
// NVRTC (loaded lazily with dlopen: the library must load on machines without it)
// ------------------------------------------------------------------------------------------------------------------
struct Nvrtc {
  typedef struct _nvrtcProgram* prog;
  int (*CreateProgram)(prog*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(prog, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(prog, size_t*) = nullptr;
  int (*GetCUBIN)(prog, char*) = nullptr;
  int (*GetProgramLogSize)(prog, size_t*) = nullptr;
  int (*GetProgramLog)(prog, char*) = nullptr;
  int (*DestroyProgram)(prog*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
  std::string why;
};

Nvrtc* nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    std::vector<std::string> names;
    if (const char* e = getenv("NBDT_B200_NVRTC")) names.push_back(e);
    for (const char* s : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so",
                          "libnvrtc.so.13"})
      names.push_back(s);
    void* h = nullptr;
    for (const std::string& s : names) {
      h = dlopen(s.c_str(), RTLD_NOW | RTLD_LOCAL);
      if (h) break;
    }
    if (!h) { n.why = "libnvrtc not found (set NBDT_B200_NVRTC to its path)"; return; }
#define NBDT_SYM(name)                                                                      \
  *(void**)(&n.name) = dlsym(h, "nvrtc" #name);                                              \
  if (!n.name) { n.why = "libnvrtc lacks nvrtc" #name; return; }
    NBDT_SYM(CreateProgram) NBDT_SYM(CompileProgram) NBDT_SYM(GetCUBINSize) NBDT_SYM(GetCUBIN)
    NBDT_SYM(GetProgramLogSize) NBDT_SYM(GetProgramLog) NBDT_SYM(DestroyProgram) NBDT_SYM(GetErrorString)
#undef NBDT_SYM
    n.ok = true;
  });
  return &n;
}

int compile_cubin(const std::string& src, std::vector<char>* cubin, std::string* log) {
  Nvrtc* n = nvrtc();
  if (!n->ok) { *log = n->why; return -1; }
  Nvrtc::prog prog = nullptr;
  int rc = n->CreateProgram(&prog, src.c_str(), "nbdt_tree_ka.cu", 0, nullptr, nullptr);
  if (rc) { *log = std::string("nvrtcCreateProgram: ") + n->GetErrorString(rc); return -1; }
  // -lineinfo quadruples the cubin (5.5 MB at 1000 leaves): only for profiling (ncu --import-source), NBDT_B200_KA_LINEINFO=1
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--ptxas-options=-v", "-default-device", "-lineinfo"};
  rc = n->CompileProgram(prog, getenv("NBDT_B200_KA_LINEINFO") ? 5 : 4, opts);
  size_t ls = 0;
  if (n->GetProgramLogSize(prog, &ls) == 0 && ls > 1) {
    log->resize(ls);
    n->GetProgramLog(prog, &(*log)[0]);
  }
  if (rc) {
    *log = std::string("nvrtcCompileProgram: ") + n->GetErrorString(rc) + "\n" + *log;
    n->DestroyProgram(&prog);
    return -1;
  }
  size_t cs = 0;
  rc = n->GetCUBINSize(prog, &cs);
  if (rc == 0 && cs > 0) {
    cubin->resize(cs);
    rc = n->GetCUBIN(prog, cubin->data());
  }
  n->DestroyProgram(&prog);
  if (rc || cs == 0) { *log = "nvrtcGetCUBIN failed"; return -1; }
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// on-disk cache
// ------------------------------------------------------------------------------------------------------------------
std::string cache_dir() {
  static std::string dir;
  static std::once_flag once;
  std::call_once(once, [] {
    std::vector<std::string> cand;
    if (const char* e = getenv("NBDT_B200_CACHE")) cand.push_back(e);
    Dl_info info;
    if (dladdr((void*)&cache_dir, &info) && info.dli_fname) {
      std::string p = info.dli_fname;
      const size_t s = p.rfind('/');
      if (s != std::string::npos) cand.push_back(p.substr(0, s) + "/tree_cache");
    }
    if (const char* h = getenv("HOME")) cand.push_back(std::string(h) + "/.cache/nbdt_b200");
    cand.push_back("/tmp/nbdt_b200_cache");
    for (const std::string& c : cand) {
      // mkdir -p (two levels are enough for the candidates above)
      const size_t s = c.rfind('/');
      if (s != std::string::npos && s > 0) mkdir(c.substr(0, s).c_str(), 0777);
      mkdir(c.c_str(), 0777);
      if (access(c.c_str(), W_OK | X_OK) == 0) { dir = c; return; }
    }
  });
  return dir;
}

// read-only lookups also consult the in-tree directory when the writable one is elsewhere
std::vector<std::string> cache_read_dirs() {
  std::vector<std::string> v;
  const std::string w = cache_dir();
  if (!w.empty()) v.push_back(w);
  Dl_info info;
  if (dladdr((void*)&cache_dir, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    const size_t s = p.rfind('/');
    if (s != std::string::npos) {
      const std::string d = p.substr(0, s) + "/tree_cache";
      if (d != w) v.push_back(d);
    }
  }
  return v;
}

std::string variant_name(const KTree& t, const Variant& v, const Knobs& k, bool dflt) {
  char b[192];
  int n = snprintf(b, sizeof b, "ka%d_%016llx_m%d_t%d_v%d_o%d%s", kGeneratorVersion, t.hash, v.mode, v.dtype, v.vec, v.ovec, v.layout ? "_nchw" : "");
  if (!dflt) snprintf(b + n, sizeof b - n, "_r%d_b%d_u%d_p%d_l%d_f%d_w%d_e%d_x%d_g%d_k%d", k.regcap, k.minb, k.unroll, k.prefetch, k.maxld, k.fence, k.warps, k.expm, k.rows, k.teams, k.lockstep);
  return b;
}

bool read_file(const std::string& path, std::vector<char>* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out->resize(n > 0 ? (size_t)n : 0);
  const bool ok = n > 0 && fread(out->data(), 1, (size_t)n, f) == (size_t)n;
  fclose(f);
  return ok;
}

void write_file_atomic(const std::string& path, const char* data, size_t n) {
  char tmp[32];
  snprintf(tmp, sizeof tmp, ".tmp%d", (int)getpid());
  const std::string t = path + tmp;
  FILE* f = fopen(t.c_str(), "wb");
  if (!f) return;
  const bool ok = fwrite(data, 1, n, f) == n;
  fclose(f);
  if (ok) rename(t.c_str(), path.c_str());
  else unlink(t.c_str());
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// per-tree state
// ------------------------------------------------------------------------------------------------------------------
struct TreeJit {
  int device, sm_count, max_smem;
  KTree tree;
  Plan plan;
  int policy;  // 0 off, 1 cache only, 2 compile on demand
  std::mutex mu;
  struct Loaded {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kern = nullptr;
    int blocks_per_sm = 1;
    bool failed = false;
  };
  std::map<int, Loaded> loaded;
  std::string last_variant;
};

namespace {

int jit_policy() {
  const char* e = getenv("NBDT_B200_TREE_JIT");
  if (!e || !*e) return 2;
  if (!strcmp(e, "0") || !strcmp(e, "off")) return 0;
  if (!strcmp(e, "cache")) return 1;
  return 2;
}

// cubin for one variant: cache hit, or generate + compile + store.  Returns 0 on success.
int obtain_cubin(const KTree& t, const Plan& plan, const Variant& v, int policy, std::vector<char>* cubin, std::string* log) {
  const std::string name = variant_name(t, v, plan.knobs, plan.default_knobs);
  for (const std::string& d : cache_read_dirs())
    if (read_file(d + "/" + name + ".cubin", cubin)) return 0;
  if (policy < 2) { *log = "not in the kernel cache and NBDT_B200_TREE_JIT=cache"; return -1; }
  Gen g(t, v, plan);
  const std::string src = g.run();
  if (compile_cubin(src, cubin, log)) return -1;
  const std::string dir = cache_dir();
  if (!dir.empty()) {
    write_file_atomic(dir + "/" + name + ".cubin", cubin->data(), cubin->size());
    write_file_atomic(dir + "/" + name + ".log", log->data(), log->size());
  }
  return 0;
}

int max_pow2_dividing(unsigned long long v, int cap) {
  int a = 1;
  while (a < cap && (v % (unsigned long long)(a * 2)) == 0) a *= 2;
  return a;
}

}  // namespace

int tree_jit_create(const nbdt_tree_desc* d, int device, int sm_count, int max_smem, TreeJit** out) {
  *out = nullptr;
  const int policy = jit_policy();
  if (policy == 0 || d->num_classes > kMaxLeaves || d->num_classes < 2) return NBDT_OK;
  TreeJit* t = new (std::nothrow) TreeJit();
  NBDT_REQUIRE(t != nullptr, "out of host memory");
  t->device = device; t->sm_count = sm_count; t->max_smem = max_smem; t->policy = policy;
  analyse(d, &t->tree);
  t->plan = make_plan(t->tree, max_smem, read_knobs());
  if (t->plan.smem_bytes > max_smem - 1024) { delete t; return NBDT_OK; }
  *out = t;
  return NBDT_OK;
}

void tree_jit_destroy(TreeJit* t) {
  if (!t) return;
  for (auto& kv : t->loaded)
    if (kv.second.lib) cudaLibraryUnload(kv.second.lib);
  delete t;
}

const char* tree_jit_last_variant(const TreeJit* t) { return t ? t->last_variant.c_str() : ""; }

int tree_jit_launch(TreeJit* t, int mode, int dtype, const SpecParams& p, void* stream) {
  if (!t || dtype < 0 || dtype > 2 || mode < 0 || mode > 2) return -1;
  const int esz = dtype == NBDT_F32 ? 4 : 2;
  const int C = t->tree.C;
  Variant v;
  v.mode = mode; v.dtype = dtype;
  // widest vector every row start of x (and of the gradient) is aligned to
  unsigned long long align = (unsigned long long)C * esz;
  align |= (unsigned long long)reinterpret_cast<uintptr_t>(p.x);
  align |= (unsigned long long)(p.ld * esz);
  if (mode != kSpecSoftFwd && p.grad) align |= (unsigned long long)reinterpret_cast<uintptr_t>(p.grad);
  v.vec = max_pow2_dividing(align | 16ull, 16);
  if (v.vec < esz) return -1;
  v.ovec = 4;
  if (mode == kSpecSoftFwd)
    v.ovec = max_pow2_dividing(((unsigned long long)C * 4) | (unsigned long long)reinterpret_cast<uintptr_t>(p.class_probs) | 16ull, 16);
  if (v.ovec < 4) return -1;
  if (p.hw > 0) {  // class-major tensors: 32 consecutive pixels of one image per block, element-wise accesses
    if (p.hw % 32 != 0 || t->plan.rows != 1) return -1;
    v.layout = 1; v.vec = esz; v.ovec = 4;
  }

  TreeJit::Loaded* L;
  {
    std::lock_guard<std::mutex> lock(t->mu);
    L = &t->loaded[v.key()];
    if (L->failed) return -1;
    if (!L->kern) {
      std::vector<char> cubin;
      std::string log;
      if (obtain_cubin(t->tree, t->plan, v, t->policy, &cubin, &log)) {
        L->failed = true;
        if (getenv("NBDT_B200_TREE_JIT_VERBOSE")) fprintf(stderr, "[nbdt_b200] tree JIT unavailable for %s: %s\n", variant_name(t->tree, v, t->plan.knobs, t->plan.default_knobs).c_str(), log.c_str());
        return -1;
      }
      cudaError_t e = cudaLibraryLoadData(&L->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
      if (e == cudaSuccess) e = cudaLibraryGetKernel(&L->kern, L->lib, "nbdt_tree_ka");
      if (e == cudaSuccess && t->plan.smem_bytes > 48 * 1024)
        e = cudaFuncSetAttribute((const void*)L->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, t->plan.smem_bytes);
      if (e == cudaSuccess) {
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)L->kern, t->plan.threads, t->plan.smem_bytes);
        L->blocks_per_sm = std::max(1, nb);
      }
      if (e != cudaSuccess) {
        cudaGetLastError();
        L->failed = true;
        L->kern = nullptr;
        if (getenv("NBDT_B200_TREE_JIT_VERBOSE")) fprintf(stderr, "[nbdt_b200] loading the JIT tree kernel failed: %s\n", cudaGetErrorString(e));
        return -1;
      }
    }
    t->last_variant = variant_name(t->tree, v, t->plan.knobs, t->plan.default_knobs);
  }
  const int rb = 32 * t->plan.rows;
  const long long rbt = (long long)rb * t->plan.teams;
  long long grid = (p.B + rbt - 1) / rbt;
  grid = std::min<long long>(grid, (long long)t->sm_count * L->blocks_per_sm);
  if (grid < 1) grid = 1;
  SpecParams sp = p;
  void* args[1] = {&sp};
  NBDT_CUDA_OK(cudaLaunchKernel((const void*)L->kern, dim3((unsigned)grid), dim3(t->plan.threads), args, (size_t)t->plan.smem_bytes, (cudaStream_t)stream));
  return NBDT_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// C ABI: source / cubin of one variant without a device (build-time cache population, tests, inspection)
// ------------------------------------------------------------------------------------------------------------------
extern "C" {

// Writes the generated CUDA source of (mode, dtype, vec) for `desc` into buf (NUL-terminated, truncated to cap);
// returns the full length, or -1 when the hierarchy is outside the generator's limits.
long long nbdt_tree_jit_source(const nbdt_tree_desc* desc, int mode, int dtype, int vec, char* buf, long long cap) {
  if (!desc || desc->num_classes > kMaxLeaves || desc->num_classes < 2 || mode < 0 || mode > 2 || dtype < 0 || dtype > 2) return -1;
  KTree t;
  analyse(desc, &t);
  const Plan plan = make_plan(t, 227 * 1024, read_knobs());
  Variant v;
  v.mode = mode; v.dtype = dtype; v.vec = vec;
  v.ovec = mode == kSpecSoftFwd ? max_pow2_dividing(((unsigned long long)t.C * 4) | 16ull, 16) : 4;
  if (vec == 0) {  // vec == 0 asks for the class-major (N, C, H, W) variant
    if (plan.rows != 1) return -1;
    v.layout = 1; v.vec = dtype == NBDT_F32 ? 4 : 2; v.ovec = 4;
  }
  Gen g(t, v, plan);
  const std::string s = g.run();
  if (buf && cap > 0) {
    const size_t n = std::min<size_t>((size_t)cap - 1, s.size());
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return (long long)s.size();
}

// Compiles (or finds in the cache) the cubin of the variant a launch with fully aligned rows would use, without touching a
// device.  Returns NBDT_OK, or NBDT_EINVAL with nbdt_last_error() set (no NVRTC, compile error, outside the limits).
int nbdt_tree_jit_precompile(const nbdt_tree_desc* desc, int mode, int dtype) {
  NBDT_REQUIRE(desc != nullptr, "NULL desc");
  NBDT_REQUIRE(desc->num_classes <= kMaxLeaves && desc->num_classes >= 2, "hierarchy outside the JIT generator's limits (%d leaves)", desc->num_classes);
  NBDT_REQUIRE(mode >= 0 && mode <= 2 && dtype >= 0 && dtype <= 2, "bad mode/dtype");
  KTree t;
  analyse(desc, &t);
  const Plan plan = make_plan(t, 227 * 1024, read_knobs());
  const int esz = dtype == NBDT_F32 ? 4 : 2;
  Variant v;
  v.mode = mode; v.dtype = dtype;
  v.vec = max_pow2_dividing(((unsigned long long)t.C * esz) | 16ull, 16);
  v.ovec = mode == kSpecSoftFwd ? max_pow2_dividing(((unsigned long long)t.C * 4) | 16ull, 16) : 4;
  std::vector<char> cubin;
  std::string log;
  if (obtain_cubin(t, plan, v, 2, &cubin, &log)) return nbdt::set_error(NBDT_EINVAL, "tree JIT: %s", log.c_str());
  return NBDT_OK;
}

}  // extern "C"
