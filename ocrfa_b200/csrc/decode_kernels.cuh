// CTC decode kernels: softmax / mask / renormalise / argmax / collapse / confidence and the prefix beam search.
//
//   softmax+mask+renorm+argmax  /root/reference/src/ocrfa/utils.py:628-638
//   greedy collapse             /root/reference/src/ocrfa/utils.py:1099-1115   (CTCLabelConverter.decode_greedy)
//   confidence                  /root/reference/src/ocrfa/utils.py:647-660, 446-447 (custom_mean)
//   prefix beam search          /root/reference/src/ocrfa/utils.py:949-1038 with :824-851, :904-946
//
// The reference does these on the host with two device<->host round trips per batch; here one CTA per text line
// keeps the whole line on chip. The beam search reproduces the reference's float32 arithmetic order (explicit
// __fmul_rn/__fadd_rn, no FMA contraction), its stable descending sort and its labeling quirks bit for bit.
#pragma once
#include "common.cuh"

namespace ocrfa {

// ---------------------------------------------------------------------------------------------
// greedy: one CTA (4 warps) per line. Each warp owns timesteps t = warp, warp+4, ...; a lane holds classes
// lane, lane+32, ... (C <= 256). Algorithmic HBM bytes per line: T*C*4 read + T*4 written (+T*C*4 if probs kept).
// ---------------------------------------------------------------------------------------------
constexpr int kGreedyThreads = 512;  // 16 warps per line: the per-row softmax chain is latency bound, rows are independent
constexpr int kMaxClassesPerLane = 8;  // C <= 256

struct GreedyParams {
  const float* logits;      // [B][T][C]
  const uint8_t* ignore;    // [C] 1 = class masked to probability 0 before renormalisation (never class 0)
  const uint8_t* collapse_ignore;  // [C] 1 = index dropped by the collapse (converter.ignore_idx; class 0 always)
  float* probs_out;         // [B][T][C] or null
  int32_t* idx_out;         // [B][T] argmax per timestep, or null
  float* maxp_out;          // [B][T] or null
  int32_t* labels_out;      // [B][T] collapsed label indices (first len[b] valid)
  int32_t* len_out;         // [B]
  double* conf_out;         // [B]
  int B, T, C;
};

__global__ void __launch_bounds__(kGreedyThreads) ctc_greedy_kernel(const GreedyParams p) {
  extern __shared__ uint8_t dsm[];
  int32_t* s_idx = reinterpret_cast<int32_t*>(dsm);           // [T]
  float* s_maxp = reinterpret_cast<float*>(s_idx + p.T);       // [T]
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = kGreedyThreads / 32;
  const float* lg = p.logits + (size_t)b * p.T * p.C;

  uint32_t ign_bits = 0;  // bit k: class lane+32k is masked
  for (int k = 0; k < kMaxClassesPerLane; ++k) {
    const int c = lane + 32 * k;
    if (c < p.C && p.ignore && p.ignore[c]) ign_bits |= 1u << k;
  }

  for (int t = warp; t < p.T; t += nwarps) {
    const float* row = lg + (size_t)t * p.C;
    float v[kMaxClassesPerLane];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kMaxClassesPerLane; ++k) {
      const int c = lane + 32 * k;
      v[k] = c < p.C ? __ldg(row + c) : -INFINITY;
      mx = fmaxf(mx, v[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxClassesPerLane; ++k) {
      v[k] = (lane + 32 * k) < p.C ? expf(v[k] - mx) : 0.f;
      sum += v[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    float msum = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxClassesPerLane; ++k) {
      v[k] = __fdiv_rn(v[k], sum);                 // softmax probability
      if ((ign_bits >> k) & 1u) v[k] = 0.f;        // preds_prob[:, :, ignore_idx] = 0
      msum += v[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
    float best = -1.f;
    int besti = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < kMaxClassesPerLane; ++k) {
      const int c = lane + 32 * k;
      if (c < p.C) {
        v[k] = __fdiv_rn(v[k], msum);              // renormalise
        if (v[k] > best) { best = v[k]; besti = c; }  // ascending c within the lane: lowest index wins ties
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if (p.probs_out) {
      float* po = p.probs_out + ((size_t)b * p.T + t) * p.C;
#pragma unroll
      for (int k = 0; k < kMaxClassesPerLane; ++k) {
        const int c = lane + 32 * k;
        if (c < p.C) po[c] = v[k];
      }
    }
    if (lane == 0) {
      s_idx[t] = besti;
      s_maxp[t] = best;
      if (p.idx_out) p.idx_out[(size_t)b * p.T + t] = besti;
      if (p.maxp_out) p.maxp_out[(size_t)b * p.T + t] = best;
    }
  }
  __syncthreads();

  if (warp == 0) {
    // collapse: keep t iff idx[t] != idx[t-1] (t = 0 always kept) and idx[t] not ignored
    int base = 0;
    int32_t* lo = p.labels_out + (size_t)b * p.T;
    for (int t0 = 0; t0 < p.T; t0 += 32) {
      const int t = t0 + lane;
      bool keep = false;
      int k = 0;
      if (t < p.T) {
        k = s_idx[t];
        const bool fresh = (t == 0) || (k != s_idx[t - 1]);
        const bool ign = (k == 0) || (p.collapse_ignore && p.collapse_ignore[k]);
        keep = fresh && !ign;
      }
      const uint32_t m = __ballot_sync(0xffffffffu, keep);
      if (keep) lo[base + __popc(m & ((1u << lane) - 1u))] = k;
      base += __popc(m);
    }
    if (lane == 0) {
      p.len_out[b] = base;
      // confidence: product of the max probabilities at non-blank timesteps, ** (2 / sqrt(n))
      float prod = 1.f;
      int n = 0;
      for (int t = 0; t < p.T; ++t)
        if (s_idx[t] != 0) { prod = __fmul_rn(prod, s_maxp[t]); ++n; }
      p.conf_out[b] = n > 0 ? pow(static_cast<double>(prod), 2.0 / sqrt(static_cast<double>(n))) : 0.0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// decision margin of the greedy decode: for every line, the smallest distance between the winning class and the
// runner-up over all timesteps, measured on the LOGITS of the classes that are not masked (utils.py:630: masked
// classes get probability 0 and can never win) and normalised by the rms of those logits in the row:
//     margin[b] = min_t (l_(1) - l_(2)) / rms_c(l)
// softmax / renormalisation are monotone, so argmax_c p = argmax_c l over the unmasked classes; a line whose margin
// exceeds twice the arithmetic error of the logits (relative to their rms) decodes to the same greedy string under
// ANY evaluation within that error - the reference's own fp32 result included. One warp per timestep, 8 per CTA.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ctc_margin_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ ignore,
                                                         float* __restrict__ margin_out, int T, int C) {
  __shared__ float s_min[8];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* lg = logits + (size_t)b * T * C;
  float line_min = INFINITY;
  for (int t = warp; t < T; t += 8) {
    const float* row = lg + (size_t)t * C;
    float m1 = -INFINITY, m2 = -INFINITY, sq = 0.f;
    int cnt = 0;
    for (int c = lane; c < C; c += 32) {
      if (ignore && ignore[c]) continue;
      const float v = __ldg(row + c);
      if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) m2 = v;
      sq = fmaf(v, v, sq);
      ++cnt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float o1 = __shfl_xor_sync(0xffffffffu, m1, o), o2 = __shfl_xor_sync(0xffffffffu, m2, o);
      // merge two (max, second) pairs: equal maxima in different lanes are two classes at the same logit -> margin 0
      const float hi = fmaxf(m1, o1), lo = fminf(m1, o1);
      m2 = fmaxf(lo, fmaxf(m2, o2));
      m1 = hi;
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    const float rms = sqrtf(sq / (float)(cnt > 0 ? cnt : 1));
    const float m = cnt >= 2 ? (m1 - m2) / fmaxf(rms, 1e-30f) : INFINITY;
    line_min = fminf(line_min, m);
  }
  if (lane == 0) s_min[warp] = line_min;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = s_min[0];
    for (int i = 1; i < 8; ++i) m = fminf(m, s_min[i]);
    margin_out[b] = m;
  }
}

// ---------------------------------------------------------------------------------------------
// prefix beam search (ctcBeamSearch, /root/reference/src/ocrfa/utils.py:949-1038): one warp per line, sequential
// over T. Labelings are never materialised inside the time loop: every step records, per surviving beam, a
// back-pointer (parent rank, edit op, class) in a trellis; the best labeling is rebuilt once at the end.
// ---------------------------------------------------------------------------------------------
struct BeamParams {
  const float* probs;    // [B][T][C] float32, already masked + renormalised (the reference's `mat`)
  int32_t* labels_out;   // [B][T]
  int32_t* len_out;      // [B]
  uint32_t* trellis;     // [B][T][bw] back-pointers: parent | op << 8 | class << 16 (global scratch, write-only in the loop)
  const uint8_t* collapse_ignore;  // [C] or null (class 0 always dropped)
  const int32_t* span_row0;  // optional [B]: item b reads rows span_row0[b] .. +span_len[b] of probs seen as [rows][C]
  const int32_t* span_len;   // optional [B]: timesteps of item b (<= T); null = every item has T rows at b*T
  int B, T, C, bw;
  int emax;              // bw * (C + 1)
  int hs;                // hash slots (power of two >= 2 * emax)
  int ht32;              // 1: 32-bit hash slots
  // candidate mode only (ctc_beam_kernel<true>, BeamState.wordsearch, utils.py:853-872): `bw` is then the WIDTH of the
  // state arrays and of a trellis row = max(keep, n_out); `keep` labelings survive every step but the last, which ranks
  // `bw` of them; the first n_out are rebuilt and collapsed into cand_labels [B][n_out][T] / cand_len [B][n_out],
  // cand_count [B] = how many exist
  int keep, n_out;
  int32_t* cand_labels;
  int32_t* cand_len;
  int32_t* cand_count;
};

enum : int { kOpSame = 0, kOpAppend = 1, kOpReplace = 2 };

__device__ __forceinline__ uint64_t beam_hash_mix(uint64_t h, int c) {
  uint64_t x = (h + static_cast<uint64_t>(c) + 1ull) * 0x9E3779B97F4A7C15ull;
  x ^= x >> 29;
  return x * 0xBF58476D1CE4E5B9ull + 0x94D049BB133111EBull;
}

struct BeamSmem {
  float* row; int16_t* cand;
  // live beams, double buffered: [2][bw]
  float *b_tot, *b_nb, *b_b; int32_t* b_len; int16_t *b_last, *b_last2; uint64_t *b_hash, *b_hprev;
  // candidate entries of the current step: [emax]. e_hash .. e_first are contiguous (28 B per entry) and are reused
  // after the time loop as the staging area of the trellis.
  uint64_t* e_hash;
  float *a_tot, *a_nb, *a_b;           // accumulated probabilities of the labeling this entry REPRESENTS
  float* x_nb;                         // this entry's own prNonBlank contribution
  float *xp_b, *xp_2;                  // [bw] per parent: the copy-beam's prBlank contribution / its "extend by blank" one
  int32_t* e_first;                    // creation order of the labeling (min member entry index)
  int16_t *e_len, *e_rep;              // labeling length; representative entry of this entry's labeling
  uint32_t* e_meta;                    // class | parent << 16 | op << 24 | flags << 26
  uint16_t* ht16;    // [hs] hash slots, 0xffff = empty (wide beams: the table must fit beside the entries)
  int32_t* ht32;     // [hs] ... or 32-bit slots (-1 = empty) when shared memory allows: native atomicCAS
  uint32_t* chain;   // [T] back-pointers of the winning path
  int16_t* raw;      // [T] its raw labeling
};
enum : int { kEfActive = 1, kEfSecond = 2 };
constexpr int kBeamEntryBytes = 8 + 4 * 3 + 4 + 4 + 2 * 2 + 4;  // 36
constexpr int kBeamRegChunks = 8;  // register-resident top-bw selection up to 32 * 8 candidate entries

__host__ __device__ inline size_t beam_smem_bytes(int C, int bw, int emax, int hs, int T, int ht32 = 0) {
  size_t n = 0;
  n += (size_t)C * 4 + (size_t)((C * 2 + 7) & ~7);
  n += (size_t)2 * bw * (4 * 3 + 4 + 2 * 2 + 8 * 2) + (size_t)bw * 8;
  n = (n + 7) & ~(size_t)7;
  n += (size_t)emax * kBeamEntryBytes;
  n = (n + 7) & ~(size_t)7;
  n += (size_t)hs * (ht32 ? 4 : 2);
  n += (size_t)T * 4 + (size_t)((T * 2 + 7) & ~7);
  return n + 64;
}

// One warp per line. Per timestep:
//   1. every (parent, candidate) entry is built IN PARALLEL (32 entries per pass): labeling hash/length, edit op,
//      its probability contribution; entries of equal labeling find one representative through a shared-memory
//      hash table and record the labeling's creation order (min entry index = the reference's dict insertion order);
//   2. the contributions are accumulated into the representatives parent by parent, in rank order - exactly the
//      reference's sequence of float32 `+=` (utils.py:1004-1026), so the sums are bit-identical;
//   3. the best `bw` labelings (prTotal descending, creation order ascending on ties = Python's stable sort) are
//      picked with warp redux operations over register-resident keys.
// kCand = false: the best labeling only (decode_beamsearch, dictionary-free word beam search). kCand = true: the ranked
// candidate list of the final beam state, for the dictionary lookup of BeamState.wordsearch.
template <bool kCand>
__global__ void __launch_bounds__(32) ctc_beam_kernel(const BeamParams p) {
  extern __shared__ __align__(16) uint8_t dsm[];
  const int lane = threadIdx.x;
  const int b = blockIdx.x;
  const int C = p.C, T = p.T, bw = p.bw, emax = p.emax;
  const int Tb = p.span_len ? min(p.span_len[b], T) : T;
  // ---- carve shared memory (8-byte fields first)
  uint8_t* q = dsm;
  BeamSmem s;
  s.e_hash = reinterpret_cast<uint64_t*>(q); q += (size_t)emax * 8;
  s.a_tot = reinterpret_cast<float*>(q); q += (size_t)emax * 4;
  s.a_nb = reinterpret_cast<float*>(q); q += (size_t)emax * 4;
  s.a_b = reinterpret_cast<float*>(q); q += (size_t)emax * 4;
  s.x_nb = reinterpret_cast<float*>(q); q += (size_t)emax * 4;
  s.e_first = reinterpret_cast<int32_t*>(q); q += (size_t)emax * 4;
  q = dsm + (((size_t)emax * 28 + 7) & ~(size_t)7);
  s.b_hash = reinterpret_cast<uint64_t*>(q); q += (size_t)2 * bw * 8;
  s.b_hprev = reinterpret_cast<uint64_t*>(q); q += (size_t)2 * bw * 8;
  s.row = reinterpret_cast<float*>(q); q += (size_t)C * 4;
  s.b_tot = reinterpret_cast<float*>(q); q += (size_t)2 * bw * 4;
  s.b_nb = reinterpret_cast<float*>(q); q += (size_t)2 * bw * 4;
  s.b_b = reinterpret_cast<float*>(q); q += (size_t)2 * bw * 4;
  s.b_len = reinterpret_cast<int32_t*>(q); q += (size_t)2 * bw * 4;
  s.xp_b = reinterpret_cast<float*>(q); q += (size_t)bw * 4;
  s.xp_2 = reinterpret_cast<float*>(q); q += (size_t)bw * 4;
  s.chain = reinterpret_cast<uint32_t*>(q); q += (size_t)T * 4;
  s.e_meta = reinterpret_cast<uint32_t*>(q); q += (size_t)emax * 4;
  s.ht32 = reinterpret_cast<int32_t*>(q);
  s.ht16 = reinterpret_cast<uint16_t*>(q); q += (size_t)p.hs * (p.ht32 ? 4 : 2);
  s.e_len = reinterpret_cast<int16_t*>(q); q += (size_t)emax * 2;
  s.e_rep = reinterpret_cast<int16_t*>(q); q += (size_t)emax * 2;
  s.b_last = reinterpret_cast<int16_t*>(q); q += (size_t)2 * bw * 2;
  s.b_last2 = reinterpret_cast<int16_t*>(q); q += (size_t)2 * bw * 2;
  s.cand = reinterpret_cast<int16_t*>(q); q += (size_t)C * 2;
  s.raw = reinterpret_cast<int16_t*>(q); q += (size_t)T * 2;

  const float thr = static_cast<float>(0.5 / static_cast<double>(C));  // np.float32(0.5 / maxC)
  const float* mat = p.probs + (p.span_row0 ? (size_t)p.span_row0[b] * C : (size_t)b * T * C);
  uint32_t* trel = p.trellis + (size_t)b * T * bw;
  const uint64_t kEmptyHash = 0x243F6A8885A308D3ull;

  int cur = 0;   // which half of the double-buffered beam state holds the parents
  int nb = 1;    // number of live beams
  if (lane == 0) {
    s.b_tot[0] = 1.f; s.b_nb[0] = 0.f; s.b_b[0] = 1.f; s.b_len[0] = 0;
    s.b_last[0] = -1; s.b_last2[0] = -1; s.b_hash[0] = kEmptyHash; s.b_hprev[0] = kEmptyHash;
  }
  // the probability row of the NEXT step is fetched while the current one is processed (C <= 256; larger class
  // counts read the row in place)
  const bool prefetch = C <= 32 * kMaxClassesPerLane;
  float nxt_v[kMaxClassesPerLane];
#pragma unroll
  for (int k = 0; k < kMaxClassesPerLane; ++k) {
    const int c = lane + 32 * k;
    nxt_v[k] = (prefetch && Tb > 0 && c < C) ? __ldg(mat + c) : 0.f;
  }
  __syncwarp();

#ifdef OCRFA_BEAM_PROF
  long long prof[6] = {0, 0, 0, 0, 0, 0};
  long long pt0 = clock64();
#define BEAM_PROF(i) do { const long long n_ = clock64(); prof[i] += n_ - pt0; pt0 = n_; } while (0)
#else
#define BEAM_PROF(i) do { } while (0)
#endif
  for (int t = 0; t < Tb; ++t) {
    // ---- a. stage the probability row, list the classes above threshold in ascending order
    int ncand = 0;
    if (prefetch) {
#pragma unroll
      for (int k = 0; k < kMaxClassesPerLane; ++k) {
        const int c = lane + 32 * k;
        if (32 * k < C) {  // uniform
          const float v = nxt_v[k];
          bool pred = false;
          if (c < C) { s.row[c] = v; pred = v >= thr; }
          const uint32_t m = __ballot_sync(0xffffffffu, pred);
          if (pred) s.cand[ncand + __popc(m & ((1u << lane) - 1u))] = static_cast<int16_t>(c);
          ncand += __popc(m);
        }
      }
    } else {
      const float* rowg = mat + (size_t)t * C;
      for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        float v = 0.f;
        bool pred = false;
        if (c < C) { v = __ldg(rowg + c); s.row[c] = v; pred = v >= thr; }
        const uint32_t m = __ballot_sync(0xffffffffu, pred);
        if (pred) s.cand[ncand + __popc(m & ((1u << lane) - 1u))] = static_cast<int16_t>(c);
        ncand += __popc(m);
      }
    }
    if (prefetch && t + 1 < Tb) {
      const float* rown = mat + (size_t)(t + 1) * C;
#pragma unroll
      for (int k = 0; k < kMaxClassesPerLane; ++k) {
        const int c = lane + 32 * k;
        nxt_v[k] = c < C ? __ldg(rown + c) : 0.f;
      }
    }
    const int per_parent = 1 + ncand;
    const int E = nb * per_parent;
    int hs = 64;
    while (hs < 2 * E) hs <<= 1;
    if (p.ht32) { for (int i = lane; i < hs; i += 32) s.ht32[i] = -1; }
    else { for (int i = lane; i < hs; i += 32) s.ht16[i] = 0xffffu; }
    __syncwarp();
    const bool blank_cand = ncand > 0 && s.cand[0] == 0;  // blank passes the threshold
    BEAM_PROF(0);

    const float* pt = s.b_tot + cur * bw; const float* pn = s.b_nb + cur * bw; const float* pb = s.b_b + cur * bw;
    const int32_t* pl = s.b_len + cur * bw; const int16_t* pla = s.b_last + cur * bw; const int16_t* pl2 = s.b_last2 + cur * bw;
    const uint64_t* ph = s.b_hash + cur * bw; const uint64_t* php = s.b_hprev + cur * bw;

    // ---- b1. build all entries. Entry index self = parent * per_parent + k is the reference's creation order
    // (k == 0: copy-beam, k >= 1: extension by cand[k-1]). "Extend by blank" of a parent that is empty or already
    // ends in blank IS the parent's own labeling: it has no entry of its own, the copy-beam entry carries it as a
    // second contribution (added right after its own, the reference's order).
    for (int e0 = 0; e0 < E; e0 += 32) {
      const int self = e0 + lane;
      bool active = self < E;
      int i = 0, k = 0;
      if (active) { i = self / per_parent; k = self - i * per_parent; }
      const float P_tot = pt[i], P_nb = pn[i], P_b = pb[i];
      const int P_len = pl[i];
      const int P_last = pla[i], P_last2 = pl2[i];
      const uint64_t P_hash = ph[i], P_hprev = php[i];
      int op = kOpSame, c = 0, nlen = P_len;
      uint64_t nhash = P_hash;
      float add_nb = 0.f, add_b = 0.f, add_tot = 0.f, add_2 = 0.f;
      bool second = false;
      if (active) {
        if (k == 0) {
          const float nbv = P_len > 0 ? __fmul_rn(P_nb, s.row[P_last]) : 0.f;
          const float blv = __fmul_rn(P_tot, s.row[0]);
          add_nb = nbv; add_b = blv; add_tot = __fadd_rn(blv, nbv);
          if (blank_cand && ((P_len == 0) || (P_last == 0))) {
            second = true;
            add_2 = __fmul_rn(s.row[0], (P_len > 0 && P_last == 0) ? P_b : P_tot);
          }
        } else {
          c = s.cand[k - 1];
          if (P_len == 0) op = (c == 0) ? kOpSame : kOpAppend;
          else if (c == 0) op = (P_last == 0) ? kOpSame : kOpAppend;
          else if (P_last == 0) op = (P_last2 == c) ? kOpAppend : kOpReplace;
          else op = kOpAppend;
          if (op == kOpSame) active = false;  // carried by the copy-beam entry
          if (op == kOpAppend) { nhash = beam_hash_mix(P_hash, c); nlen = P_len + 1; }
          else if (op == kOpReplace) { nhash = beam_hash_mix(P_hprev, c); nlen = P_len; }
          const float pr = __fmul_rn(s.row[c], (P_len > 0 && P_last == c) ? P_b : P_tot);
          add_nb = pr; add_tot = pr;
        }
      }
      if (self < E) {
        s.e_hash[self] = nhash; s.e_len[self] = static_cast<int16_t>(nlen);
        s.e_meta[self] = static_cast<uint32_t>(c) | (static_cast<uint32_t>(i) << 16) | (static_cast<uint32_t>(op) << 24) |
                         (static_cast<uint32_t>((active ? kEfActive : 0) | (second ? kEfSecond : 0)) << 26);
        s.x_nb[self] = add_nb;
        if (k == 0) { s.xp_b[i] = add_b; s.xp_2[i] = add_2; }
        s.a_tot[self] = 0.f; s.a_nb[self] = 0.f; s.a_b[self] = 0.f;
        s.e_first[self] = 0x7fffffff; s.e_rep[self] = -1;
      }
    }
    __syncwarp();
    // second pass: find or create the representative of every active entry's labeling
    for (int e0 = 0; e0 < E; e0 += 32) {
      const int self = e0 + lane;
      if (self < E && ((s.e_meta[self] >> 26) & kEfActive)) {
        const uint64_t nhash = s.e_hash[self];
        const int nlen = s.e_len[self];
        uint32_t slot = static_cast<uint32_t>(nhash ^ (nhash >> 32)) & (hs - 1);
        int rep;
        while (true) {
          int seen;
          if (p.ht32) {
            seen = atomicCAS(&s.ht32[slot], -1, self);
          } else {
            const unsigned short got = atomicCAS(&s.ht16[slot], static_cast<unsigned short>(0xffffu), static_cast<unsigned short>(self));
            seen = got == 0xffffu ? -1 : static_cast<int>(got);
          }
          if (seen == -1) { rep = self; break; }
          if (s.e_hash[seen] == nhash && s.e_len[seen] == nlen) { rep = seen; break; }
          slot = (slot + 1) & (hs - 1);
        }
        s.e_rep[self] = static_cast<int16_t>(rep);
        atomicMin(&s.e_first[rep], self);
      }
    }
    __syncwarp();

    BEAM_PROF(1);
    // ---- b2. ordered accumulation: parent by parent (no two entries of one parent share a labeling)
    for (int i = 0; i < nb; ++i) {
      for (int k = lane; k < per_parent; k += 32) {
        const int self = i * per_parent + k;
        const int fl = static_cast<int>(s.e_meta[self] >> 26);
        if (fl & kEfActive) {
          const int r = s.e_rep[self];
          const float xnb = s.x_nb[self];
          float vnb = __fadd_rn(s.a_nb[r], xnb);
          // prTotal contribution: the copy-beam adds (prBlank + prNonBlank), an extension its prNonBlank
          float vtot = __fadd_rn(s.a_tot[r], k == 0 ? __fadd_rn(s.xp_b[i], xnb) : xnb);
          if (k == 0) s.a_b[r] = __fadd_rn(s.a_b[r], s.xp_b[i]);
          if (fl & kEfSecond) { vnb = __fadd_rn(vnb, s.xp_2[i]); vtot = __fadd_rn(vtot, s.xp_2[i]); }
          s.a_nb[r] = vnb; s.a_tot[r] = vtot;
        }
      }
      __syncwarp();
    }

    BEAM_PROF(2);
    // ---- c. keep the best `bw` labelings: prTotal descending, creation order ascending on ties (stable sort)
    const int nxt = cur ^ 1;
    uint32_t* trow = trel + (size_t)t * bw;
    int n_new = 0;
    // labelings ranked by this step: all `bw` of them, except that candidate mode keeps only `keep` before the last step
    const int sel = (kCand && t != Tb - 1) ? p.keep : bw;
    auto emit = [&](int j, int r) {
      // labeling represented by entry j becomes beam r of the next step
      const uint32_t meta = s.e_meta[j];
      const int par = (meta >> 16) & 0xff, op = (meta >> 24) & 3, c = meta & 0xffff;
      const int o = nxt * bw + r;
      s.b_tot[o] = s.a_tot[j]; s.b_nb[o] = s.a_nb[j]; s.b_b[o] = s.a_b[j]; s.b_len[o] = s.e_len[j];
      s.b_hash[o] = s.e_hash[j];
      if (op == kOpSame) { s.b_last[o] = pla[par]; s.b_last2[o] = pl2[par]; s.b_hprev[o] = php[par]; }
      else if (op == kOpAppend) { s.b_last[o] = static_cast<int16_t>(c); s.b_last2[o] = pla[par]; s.b_hprev[o] = ph[par]; }
      else { s.b_last[o] = static_cast<int16_t>(c); s.b_last2[o] = pl2[par]; s.b_hprev[o] = php[par]; }
      trow[r] = static_cast<uint32_t>(par) | (static_cast<uint32_t>(op) << 8) | (static_cast<uint32_t>(c) << 16);
    };
    // key of a representative: float bits of prTotal (non-negative, so integer order = float order) + 1; 0 = not a candidate
    if (E <= 32 * kBeamRegChunks) {
      uint32_t key[kBeamRegChunks];
      int first[kBeamRegChunks];
#pragma unroll
      for (int m = 0; m < kBeamRegChunks; ++m) {
        const int j = lane + 32 * m;
        key[m] = 0u; first[m] = 0x7fffffff;
        if (j < E && s.e_rep[j] == j) { key[m] = __float_as_uint(s.a_tot[j]) + 1u; first[m] = s.e_first[j]; }
      }
      // the rounds only pick the winners (two warp reductions each); the state of the surviving beams is written
      // afterwards by bw lanes in parallel instead of one lane per round
      int my_j0 = -1, my_j1 = -1;  // winners of rounds lane and lane + 32 (bw <= 64)
      for (int r = 0; r < sel; ++r) {
        uint32_t lmax = 0u;
#pragma unroll
        for (int m = 0; m < kBeamRegChunks; ++m) lmax = key[m] > lmax ? key[m] : lmax;
        const uint32_t wmax = __reduce_max_sync(0xffffffffu, lmax);
        if (wmax == 0u) break;
        int lfirst = 0x7fffffff;
#pragma unroll
        for (int m = 0; m < kBeamRegChunks; ++m) lfirst = (key[m] == wmax && first[m] < lfirst) ? first[m] : lfirst;
        const int wfirst = __reduce_min_sync(0xffffffffu, lfirst);
        int lj = -1;
#pragma unroll
        for (int m = 0; m < kBeamRegChunks; ++m)
          if (key[m] == wmax && first[m] == wfirst) { lj = lane + 32 * m; key[m] = 0u; }
        const int wj = __reduce_max_sync(0xffffffffu, lj);  // exactly one lane holds the winner (creation order is unique)
        if ((r & 31) == lane) { if (r < 32) my_j0 = wj; else my_j1 = wj; }
        ++n_new;
      }
      if (my_j0 >= 0) emit(my_j0, lane);
      if (my_j1 >= 0) emit(my_j1, lane + 32);
    } else {
      // many candidates (flat distributions): scan shared memory each round
      for (int r = 0; r < sel; ++r) {
        uint32_t lmax = 0u;
        for (int j = lane; j < E; j += 32)
          if (s.e_rep[j] == j) { const uint32_t kx = __float_as_uint(s.a_tot[j]) + 1u; lmax = kx > lmax ? kx : lmax; }
        const uint32_t wmax = __reduce_max_sync(0xffffffffu, lmax);
        if (wmax == 0u) break;
        int lfirst = 0x7fffffff, lj = -1;
        for (int j = lane; j < E; j += 32)
          if (s.e_rep[j] == j && __float_as_uint(s.a_tot[j]) + 1u == wmax && s.e_first[j] < lfirst) { lfirst = s.e_first[j]; lj = j; }
        const int wfirst = __reduce_min_sync(0xffffffffu, lfirst);
        if (lj >= 0 && lfirst == wfirst) { emit(lj, r); s.e_rep[lj] = -1; }
        __syncwarp();
        ++n_new;
      }
    }
    nb = n_new;
    cur = nxt;
    __syncwarp();
    BEAM_PROF(3);
  }
#ifdef OCRFA_BEAM_PROF
  if (b == 0 && lane == 0)
    printf("[beam prof] T=%d cycles/step: stage+clear %lld, build %lld, accumulate %lld, select %lld\n", Tb,
           prof[0] / Tb, prof[1] / Tb, prof[2] / Tb, prof[3] / Tb);
#endif

  // ---- rebuild the best labeling from the trellis, then drop ignored indices and adjacent duplicates of the RAW
  // labeling (utils.py:1031-1035). The trellis of this line is staged in the (now dead) entry arrays when it fits.
  const size_t trel_words = (size_t)Tb * bw;
  const uint32_t* tsrc = trel;
  if (trel_words * 4 <= (size_t)emax * 28) {
    uint32_t* stg = reinterpret_cast<uint32_t*>(s.e_hash);
    for (size_t i = lane; i < trel_words; i += 32) stg[i] = trel[i];
    tsrc = stg;
  }
  __syncwarp();
  if constexpr (!kCand) {
  if (lane == 0) {
    int len = 0;
    if (nb > 0) {
      int r = 0;
      for (int t = Tb - 1; t >= 0; --t) {
        const uint32_t e = tsrc[(size_t)t * bw + r];
        s.chain[t] = e;
        r = static_cast<int>(e & 0xffu);
      }
      for (int t = 0; t < Tb; ++t) {
        const uint32_t e = s.chain[t];
        const int op = (e >> 8) & 0xff, c = static_cast<int>(e >> 16);
        if (op == kOpAppend) s.raw[len++] = static_cast<int16_t>(c);
        else if (op == kOpReplace) s.raw[len - 1] = static_cast<int16_t>(c);
      }
    }
    int32_t* lo = p.labels_out + (size_t)b * T;
    int n = 0;
    for (int i = 0; i < len; ++i) {
      const int l = s.raw[i];
      const bool ign = (l == 0) || (p.collapse_ignore && p.collapse_ignore[l]);
      if (!ign && !(i > 0 && s.raw[i - 1] == l)) lo[n++] = l;
    }
    p.len_out[b] = n;
  }
  } else {
    // ranks 0 .. n_cand-1 of the final state, best first (sortedBeams[:maxCandidate], utils.py:855-857); with no
    // timestep at all the state is the initial empty labeling
    const int n_cand = min(nb, p.n_out);
    if (lane == 0) {
      p.cand_count[b] = n_cand;
      for (int r0 = 0; r0 < n_cand; ++r0) {
        int len = 0;
        int r = r0;
        for (int t = Tb - 1; t >= 0; --t) {
          const uint32_t e = tsrc[(size_t)t * bw + r];
          s.chain[t] = e;
          r = static_cast<int>(e & 0xffu);
        }
        for (int t = 0; t < Tb; ++t) {
          const uint32_t e = s.chain[t];
          const int op = (e >> 8) & 0xff, c = static_cast<int>(e >> 16);
          if (op == kOpAppend) s.raw[len++] = static_cast<int16_t>(c);
          else if (op == kOpReplace) s.raw[len - 1] = static_cast<int16_t>(c);
        }
        int32_t* lo = p.cand_labels + ((size_t)b * p.n_out + r0) * T;
        int n = 0;
        for (int i = 0; i < len; ++i) {
          const int l = s.raw[i];
          const bool ign = (l == 0) || (p.collapse_ignore && p.collapse_ignore[l]);
          if (!ign && !(i > 0 && s.raw[i - 1] == l)) lo[n++] = l;
        }
        p.cand_len[(size_t)b * p.n_out + r0] = n;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// stand-alone collapse (CTCLabelConverter.decode_greedy on caller-provided indices, utils.py:1099-1115):
// one warp per line, ragged lines given by offsets[B+1].
// ---------------------------------------------------------------------------------------------
__global__ void ctc_collapse_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ offsets, int B, int C,
                                    const uint8_t* __restrict__ collapse_ignore, int32_t* __restrict__ labels,
                                    int32_t* __restrict__ len) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int o0 = offsets[b], o1 = offsets[b + 1];
  int base = 0;
  for (int t0 = o0; t0 < o1; t0 += 32) {
    const int t = t0 + lane;
    bool keep = false;
    int k = 0;
    if (t < o1) {
      k = idx[t];
      const bool fresh = (t == o0) || (k != idx[t - 1]);
      const bool ign = (k == 0) || (collapse_ignore && k > 0 && k < C && collapse_ignore[k]);
      keep = fresh && !ign;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (keep) labels[o0 + base + __popc(m & ((1u << lane) - 1u))] = k;
    base += __popc(m);
  }
  if (lane == 0) len[b] = base;
}

// argmax over the class axis of a [rows][C] float matrix, first maximum wins (np.argmax, utils.py:1126)
__global__ void argmax_rows_kernel(const float* __restrict__ mat, long long rows, int C, int32_t* __restrict__ out) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* row = mat + r * C;
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float v = __ldg(row + c);
    if (v > best || besti == 0x7fffffff) { best = v; besti = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (oi != 0x7fffffff && (besti == 0x7fffffff || ob > best || (ob == best && oi < besti))) { best = ob; besti = oi; }
  }
  if (lane == 0) out[r] = besti;
}

}  // namespace ocrfa
