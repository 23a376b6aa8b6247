// Interaction kernel of a MIRRORED supergroup (companies: primary activity, static members, contact matrix of
// dimension <= 4): Interaction.time_step_for_group (interaction.py:516-641) over a slot-ordered copy of the
// members' state bytes.  Included by jb_kernels.cuh after jb_groups.cuh.
//
// Everything a slot needs is read from arrays laid out in slot order -- the state byte (`marr`, kept current by
// the placement), the slot descriptor (`tp_info`), the member's person index (`tp_member`) and global id
// (`mgid`) -- so every load of the kernel is a coalesced stream; the only gathers left are the 8 bytes of an
// infector's transmission probability and of a susceptibility that is neither 0 nor 1.
//
// One WARP per work item, lane == slot of a 32-slot tile, no block-level barrier:
//   * a span of up to JB_SPAN_TILES consecutive tiles of small groups (<= 32 slots, never split by a tile):
//     presence / infector / group-boundary masks come from three ballots, a lane finds its own group with a few
//     bit operations, the infectors' transmission probabilities are summed in slot order (the reference's
//     left-to-right Python sum, interaction.py:463-467) by a loop over the set bits of the infector mask;
//   * one group of more than 32 slots (it owns whole tiles): first pass subgroup sizes and sums, second pass
//     the susceptibles; groups that fit JB_SPAN_TILES tiles stay in registers between the passes.
// The Bernoulli draws (Philox + exp, the expensive part) are queued per warp and executed 32 at a time on fully
// populated warps, as in k_tiles.
#pragma once

#define JB_SPAN_TILES 8
#define JB_SPAN_THREADS 128
#define JB_SPAN_WARPS (JB_SPAN_THREADS / 32)
#define JB_SPAN_QUEUE 64
#define JB_SPAN_SMALL 0x80000000u  // item.y: bit 31 = span of small-group tiles (low bits: tiles), else a group index

__device__ __forceinline__ uint32_t jb_mb_var(uint32_t mb) { return ((mb >> 2) & 1u) | ((mb >> 6) & 2u); }

// Rare path (a draw succeeded and events are recorded): _blame_subgroup / _blame_individuals
// (interaction.py:643-662) for the member in slot x of the group that owns slots [s0, s0 + len).
template <int VM>
__device__ __noinline__ uint32_t jb_blame_span(const GroupArgs& a, const double* trans, uint32_t absmask, uint32_t s0,
                                               uint32_t len, uint32_t x, uint32_t g, int v, uint32_t p) {
  const uint8_t* marr = a.marr + a.stream_base;
  const int dim = a.dim;
  uint32_t npk[4] = {0u, 0u, 0u, 0u};
  double Tj[4] = {0.0, 0.0, 0.0, 0.0};
  uint32_t my_i = 0;
  for (uint32_t m = 0; m < len; ++m) {
    const uint32_t s = s0 + m, mb = marr[s];
    if (mb & absmask) continue;
    const uint32_t l = dim > 1 ? JB_TI_LSG(a.tp_info[s]) : 0u;
    if (s == x) my_i = l;
    const bool inf = (mb & 0x08u) && (VM == 1 || (int)jb_mb_var(mb) == v);
    const double t = inf ? trans[a.tp_member[s]] : 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((uint32_t)j == l) { npk[j] += 1u; if (inf) Tj[j] += t; }
  }
  const uint32_t gi = a.grp_info[g];
  const double beta = a.beta[JB_GINFO_BETA(gi) * a.R + JB_GINFO_REGION(gi)] * a.beta_scale[JB_GINFO_SCALE(gi)];
  const double dt = a.sp->dt;
  double term[4], total = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    term[j] = 0.0;
    if (j < dim && Tj[j] != 0.0) {
      const int nj = (int)npk[j];
      const int nn = ((int)my_i == j) ? max(1, nj - 1) : nj;
      term[j] = a.cm[my_i * dim + j] * Tj[j] / (double)nn * beta * dt;
      total += term[j];
    }
  }
  double u_sub, u_ind;
  jb_blame_uniforms(a, p, u_sub, u_ind);
  const double target = u_sub * total;
  double c = 0.0;
  int jb = -1, jlast = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (term[j] != 0.0 && jb < 0) { c += term[j]; jlast = j; if (c > target) jb = j; }
  if (jb < 0) jb = jlast;
  double tsel = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j == jb) tsel = Tj[j];
  const double target2 = u_ind * tsel;
  double c2 = 0.0;
  uint32_t infector = 0xFFFFFFFFu;
  for (uint32_t m = 0; m < len; ++m) {
    const uint32_t s = s0 + m, mb = marr[s];
    if ((mb & absmask) || !(mb & 0x08u)) continue;
    if ((dim > 1 ? (int)JB_TI_LSG(a.tp_info[s]) : 0) != jb) continue;
    if (VM > 1 && (int)jb_mb_var(mb) != v) continue;
    const uint32_t q = a.tp_member[s];
    c2 += trans[q];
    infector = q;
    if (c2 > target2) break;
  }
  return infector;
}

// The draw queue of one warp (shared memory) and its two out-of-line operations.  They are real functions on
// purpose: Philox + exp + the rare blame path are a few hundred instructions, and one copy of them (instead of one
// per call site of an unrolled tile loop) is what keeps the kernel inside the instruction cache.
template <int VM>
struct JbSpanQ {
  uint32_t* x;     // [Q] slot of the susceptible (relative to the supergroup's first slot)
  uint32_t* ord;   // [Q] ordinal of the injected uniform (test mode)
  double* lam;     // [VM][Q]
  uint32_t *g, *s0, *len;  // [Q] blame context (events only): group, its first slot, its slots
};

template <int VM, bool EV>
__device__ __noinline__ void jb_span_drain(const GroupArgs& a, const JbSpanQ<VM>& q, uint32_t n, uint32_t absmask) {
  const int lane = threadIdx.x & 31;
  if ((uint32_t)lane < n) {
    double lam_v[VM];
    double lam = 0.0;
#pragma unroll
    for (int v = 0; v < VM; ++v) { lam_v[v] = q.lam[v * JB_SPAN_QUEUE + lane]; lam += lam_v[v]; }
    const uint32_t x = q.x[lane];
    const int v = jb_draw_core<VM>(a, a.inject ? 0u : a.mgid[a.stream_base + x], lam_v, lam, q.ord[lane]);
    if (v >= 0) {
      const uint32_t p = a.tp_member[x];
      uint32_t infector = 0xFFFFFFFFu, g = 0;
      if (EV) {
        g = q.g[lane];
        infector = jb_blame_span<VM>(a, jb_trans_read(a.trans, a.NP, a.sp), absmask, q.s0[lane], q.len[lane], x, g, v, p);
      }
      jb_emit(a, p, (uint32_t)v, g, infector);
    }
  }
  __syncwarp();
}

// 32 or more draws are waiting: execute the first 32, move the rest to the front.  Returns the new length.
template <int VM, bool EV>
__device__ __noinline__ uint32_t jb_span_flush(const GroupArgs& a, const JbSpanQ<VM>& q, uint32_t qn, uint32_t absmask) {
  const int lane = threadIdx.x & 31;
  jb_span_drain<VM, EV>(a, q, 32u, absmask);
  const uint32_t rest = qn - 32u;
  uint32_t mx = 0, mo = 0, mg = 0, ms = 0, ml_ = 0;
  double ml[VM];
  if ((uint32_t)lane < rest) {
    mx = q.x[32 + lane]; mo = q.ord[32 + lane];
    if (EV) { mg = q.g[32 + lane]; ms = q.s0[32 + lane]; ml_ = q.len[32 + lane]; }
#pragma unroll
    for (int v = 0; v < VM; ++v) ml[v] = q.lam[v * JB_SPAN_QUEUE + 32 + lane];
  }
  __syncwarp();
  if ((uint32_t)lane < rest) {
    q.x[lane] = mx; q.ord[lane] = mo;
    if (EV) { q.g[lane] = mg; q.s0[lane] = ms; q.len[lane] = ml_; }
#pragma unroll
    for (int v = 0; v < VM; ++v) q.lam[v * JB_SPAN_QUEUE + lane] = ml[v];
  }
  __syncwarp();
  return rest;
}

template <int VM, bool EV>
__global__ void __launch_bounds__(JB_SPAN_THREADS, VM == 1 && !EV ? 8 : 1) k_span(const __grid_constant__ GroupArgs a) {
  jb_stamp(3); jb_stamp(4);
  __shared__ uint32_t q_x[JB_SPAN_WARPS][JB_SPAN_QUEUE], q_ord[JB_SPAN_WARPS][JB_SPAN_QUEUE];
  __shared__ double q_lam[JB_SPAN_WARPS][VM][JB_SPAN_QUEUE];
  __shared__ uint32_t q_g[EV ? JB_SPAN_WARPS : 1][EV ? JB_SPAN_QUEUE : 1], q_s0[EV ? JB_SPAN_WARPS : 1][EV ? JB_SPAN_QUEUE : 1];
  __shared__ uint32_t q_len[EV ? JB_SPAN_WARPS : 1][EV ? JB_SPAN_QUEUE : 1];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t item_idx = blockIdx.x * JB_SPAN_WARPS + wib;
  if (item_idx >= a.n_span) return;
  const uint2 item = a.span_items[item_idx];
  const double* const trans = jb_trans_read(a.trans, a.NP, a.sp);
  const uint32_t absmask = jb_mirror_absmask(a.sp->activity_mask, a.sp->flags);
  const uint8_t* const marr = a.marr + a.stream_base;
  const int dim = a.dim;
  const int V = (VM > 1) ? a.V : 1;
  const double dt = a.sp->dt;
  const uint32_t lane_lt = (1u << lane) - 1u;
  const bool need_p = VM > 1 || a.lambda != nullptr;  // the person index of every susceptible (else: only code READ)
  JbSpanQ<VM> q;
  q.x = q_x[wib]; q.ord = q_ord[wib]; q.lam = &q_lam[wib][0][0];
  q.g = q_g[EV ? wib : 0]; q.s0 = q_s0[EV ? wib : 0]; q.len = q_len[EV ? wib : 0];
  uint32_t qn = 0;     // warp-uniform queue length
  uint32_t draws = 0;  // warp-uniform

  // queue the draws that can succeed (lambda == 0 never infects); executed when 32 are waiting
  auto enqueue = [&](bool want, uint32_t x, const double* lam_v, uint32_t ord, uint32_t g, uint32_t s0, uint32_t len) {
    const uint32_t wm = __ballot_sync(JB_FULL, want);
    if (!wm) return;
    if (want) {
      const uint32_t pos = qn + __popc(wm & lane_lt);
      q.x[pos] = x;
      q.ord[pos] = ord;
      if (EV) { q.g[pos] = g; q.s0[pos] = s0; q.len[pos] = len; }
#pragma unroll
      for (int v = 0; v < VM; ++v) q.lam[v * JB_SPAN_QUEUE + pos] = lam_v[v];
    }
    qn += __popc(wm);
    __syncwarp();
    if (qn >= 32) qn = jb_span_flush<VM, EV>(a, q, qn, absmask);
  };
  // Lambda of a susceptible (interaction.py:594-631) given the summed transmission probabilities and sizes of
  // the subgroups of its group; returns the sum over the variants
  auto lambda_of = [&](uint32_t mb, uint32_t slot, int i, const double (&T)[VM][4], const uint32_t (&nj)[4], double beta,
                       double* lam_v) -> double {
    const uint32_t code = mb & 0x03u;
    const bool rd = need_p || code == JB_SC_READ;
    const uint32_t p = rd ? a.tp_member[slot] : 0u;
    double lam = 0.0;
#pragma unroll
    for (int v = 0; v < VM; ++v) {
      double r = 0.0;
      if (v < V) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < dim && T[v][j] != 0.0) {
            const int n_ = (int)nj[j];
            const int nn = (i == j) ? max(1, n_ - 1) : n_;  // interaction.py:469-471
            r += a.cm[i * dim + j] * T[v][j] / (double)nn * beta * dt;
          }
        r *= (VM == 1) ? (code == JB_SC_ONE ? 1.0 : code == JB_SC_ZERO ? 0.0 : a.susc[p]) : a.susc[(size_t)v * a.NP + p];
      }
      lam_v[v] = r;
      lam += r;
    }
    if (a.lambda) a.lambda[p] = lam;
    return lam;
  };

  const uint32_t slot0 = item.x;
  if (item.y & JB_SPAN_SMALL) {
    // ---- a span of tiles of small groups ---------------------------------------------------------------------
    // the span's state bytes and slot descriptors: byte k of the two 64-bit words belongs to tile k
    const uint32_t nt = item.y & 0xFFu;
    unsigned long long mb8 = 0ull, info8 = 0ull;
#pragma unroll
    for (int k = 0; k < JB_SPAN_TILES; ++k) {
      const bool ok = (uint32_t)k < nt;
      mb8 |= (unsigned long long)(ok ? (uint32_t)marr[slot0 + 32 * k + lane] : JB_MB_ABSENT) << (8 * k);
      info8 |= (unsigned long long)(ok ? (uint32_t)a.tp_info[slot0 + 32 * k + lane] : 0u) << (8 * k);
    }
    const uint32_t tg0_l = (uint32_t)lane < nt ? a.tile_group0[(slot0 >> 5) + lane] : 0u;
#pragma unroll 1
    for (uint32_t k = 0; k < nt; ++k) {
      const uint32_t mb = (uint32_t)(mb8 >> (8 * k)) & 0xFFu, info = (uint32_t)(info8 >> (8 * k)) & 0xFFu;
      const bool pres = (info & JB_TI_VALID) && !(mb & absmask);
      const bool inf = pres && (mb & 0x08u);
      const uint32_t pres32 = __ballot_sync(JB_FULL, pres), inf32 = __ballot_sync(JB_FULL, inf);
      if (!inf32) continue;
      const uint32_t first32 = __ballot_sync(JB_FULL, (info & JB_TI_FIRST) != 0u);
      const uint32_t le = 0xFFFFFFFFu >> (31 - lane);
      const uint32_t below = first32 & le, above = first32 & ~le;
      const uint32_t start = 31 - __clz(below | 1u);
      const uint32_t end = above ? (uint32_t)__ffs(above) - 1u : 32u;
      const uint32_t gmask = (end == 32u ? 0xFFFFFFFFu : ((1u << end) - 1u)) & ~((1u << start) - 1u);
      const uint32_t sus32 = pres32 & ~inf32;
      const bool has_inf = (gmask & inf32) != 0u;
      const bool act = pres && !inf && has_inf;  // a susceptible of a group that must timestep (interactive.py:136)
      const uint32_t g = __shfl_sync(JB_FULL, tg0_l, k) + __popc(below) - 1u;
      if (a.mode == 1) {  // only count the draws per group (injected-uniform ordinals)
        if ((info & JB_TI_FIRST) && has_inf && (gmask & sus32)) a.draw_cnt[g] = __popc(gmask & sus32);
        continue;
      }
      const uint32_t act32 = __ballot_sync(JB_FULL, act);
      if (!act32) continue;
      draws += __popc(act32);
      const uint32_t slot = slot0 + 32 * k + lane;
      // transmission probabilities of the infectors that have somebody to infect, summed per subgroup in slot order
      const bool need_t = inf && (gmask & sus32) != 0u;
      const double tv = need_t ? trans[a.tp_member[slot]] : 0.0;
      const uint32_t lsg = dim > 1 ? JB_TI_LSG(info) : 0u;
      const uint32_t var = VM > 1 ? jb_mb_var(mb) : 0u;
      double T[VM][4];
#pragma unroll
      for (int v = 0; v < VM; ++v)
#pragma unroll
        for (int j = 0; j < 4; ++j) T[v][j] = 0.0;
      for (uint32_t rem = inf32; rem; rem &= rem - 1u) {
        const int kb = __ffs(rem) - 1;
        const double tk = __shfl_sync(JB_FULL, tv, kb);
        const uint32_t lk = dim > 1 ? __shfl_sync(JB_FULL, lsg, kb) : 0u;
        const uint32_t vk = VM > 1 ? __shfl_sync(JB_FULL, var, kb) : 0u;
        if ((gmask >> kb) & 1u) {
#pragma unroll
          for (int v = 0; v < VM; ++v)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if ((uint32_t)v == vk && (uint32_t)j == lk) T[v][j] += tk;
        }
      }
      // subgroup sizes of the lane's group
      uint32_t nj[4] = {(uint32_t)__popc(gmask & pres32), 0u, 0u, 0u};
      if (dim > 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t sub = __ballot_sync(JB_FULL, lsg == (uint32_t)j);
          nj[j] = __popc(gmask & pres32 & sub);
        }
      }
      double lam_v[VM];
      double lam = 0.0;
#pragma unroll
      for (int v = 0; v < VM; ++v) lam_v[v] = 0.0;
      if (act) {
        const uint32_t gi = a.grp_info[g];
        const double beta = a.beta[JB_GINFO_BETA(gi) * a.R + JB_GINFO_REGION(gi)] * a.beta_scale[JB_GINFO_SCALE(gi)];
        lam = lambda_of(mb, slot, (int)lsg, T, nj, beta, lam_v);
      }
      const uint32_t ord = a.inject ? a.draw_base[g] + __popc(gmask & sus32 & lane_lt) : 0u;
      uint32_t s0 = 0, len = 0;
      if (EV) { s0 = slot0 + 32 * k + start; len = __popc(gmask & __ballot_sync(JB_FULL, (info & JB_TI_VALID) != 0u)); }
      enqueue(act && lam != 0.0, slot, lam_v, ord, g, s0, len);
    }
  } else {
    // ---- one group of more than 32 slots -------------------------------------------------------------------------
    const uint32_t g = item.y;
    const uint32_t n = a.grp_off[g + 1] - a.grp_off[g];
    const uint32_t ntile = (n + 31u) >> 5;
    unsigned long long mb8 = 0ull, info8 = 0ull;
    auto load_chunk = [&](uint32_t c0) {
      mb8 = 0ull; info8 = 0ull;
#pragma unroll
      for (int k = 0; k < JB_SPAN_TILES; ++k) {
        const uint32_t idx = (c0 + k) * 32 + lane;
        const bool ok = idx < n;
        mb8 |= (unsigned long long)(ok ? (uint32_t)marr[slot0 + idx] : JB_MB_ABSENT) << (8 * k);
        if (dim > 1) info8 |= (unsigned long long)(ok ? (uint32_t)a.tp_info[slot0 + idx] : 0u) << (8 * k);
      }
    };
    // pass 1: subgroup sizes, susceptibles, summed transmission probabilities (slot order)
    uint32_t nj[4] = {0u, 0u, 0u, 0u}, nsus = 0, ninf = 0;
    double T[VM][4];
#pragma unroll
    for (int v = 0; v < VM; ++v)
#pragma unroll
      for (int j = 0; j < 4; ++j) T[v][j] = 0.0;
    for (uint32_t c0 = 0; c0 < ntile; c0 += JB_SPAN_TILES) {
      load_chunk(c0);
      const uint32_t kn = min((uint32_t)JB_SPAN_TILES, ntile - c0);
#pragma unroll 1
      for (uint32_t k = 0; k < kn; ++k) {
        const uint32_t mb = (uint32_t)(mb8 >> (8 * k)) & 0xFFu;
        const bool pres = !(mb & absmask);
        const bool inf = pres && (mb & 0x08u);
        const uint32_t pres32 = __ballot_sync(JB_FULL, pres), inf32 = __ballot_sync(JB_FULL, inf);
        const uint32_t lsg = dim > 1 ? JB_TI_LSG((uint32_t)(info8 >> (8 * k)) & 0xFFu) : 0u;
        if (dim > 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j) nj[j] += __popc(pres32 & __ballot_sync(JB_FULL, lsg == (uint32_t)j));
        } else {
          nj[0] += __popc(pres32);
        }
        nsus += __popc(pres32 & ~inf32);
        ninf += __popc(inf32);
        if (inf32) {
          const double tv = inf ? trans[a.tp_member[slot0 + (c0 + k) * 32 + lane]] : 0.0;
          const uint32_t var = VM > 1 ? jb_mb_var(mb) : 0u;
          for (uint32_t rem = inf32; rem; rem &= rem - 1u) {
            const int kb = __ffs(rem) - 1;
            const double tk = __shfl_sync(JB_FULL, tv, kb);
            const uint32_t lk = dim > 1 ? __shfl_sync(JB_FULL, lsg, kb) : 0u;
            const uint32_t vk = VM > 1 ? __shfl_sync(JB_FULL, var, kb) : 0u;
#pragma unroll
            for (int v = 0; v < VM; ++v)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if ((uint32_t)v == vk && (uint32_t)j == lk) T[v][j] += tk;
          }
        }
      }
    }
    const bool must = nsus > 0 && ninf > 0;  // interactive.py:136
    if (a.mode == 1) {
      if (lane == 0) a.draw_cnt[g] = must ? nsus : 0u;
      return;
    }
    if (!must) return;
    draws += nsus;
    const uint32_t gi = a.grp_info[g];
    const double beta = a.beta[JB_GINFO_BETA(gi) * a.R + JB_GINFO_REGION(gi)] * a.beta_scale[JB_GINFO_SCALE(gi)];
    const uint32_t ord_base = a.inject ? a.draw_base[g] : 0u;
    // pass 2: the susceptibles (a group of <= JB_SPAN_TILES tiles is still in registers)
    uint32_t running = 0;
    for (uint32_t c0 = 0; c0 < ntile; c0 += JB_SPAN_TILES) {
      if (ntile > JB_SPAN_TILES) load_chunk(c0);
      const uint32_t kn = min((uint32_t)JB_SPAN_TILES, ntile - c0);
#pragma unroll 1
      for (uint32_t k = 0; k < kn; ++k) {
        const uint32_t mb = (uint32_t)(mb8 >> (8 * k)) & 0xFFu;
        const bool sus = !(mb & absmask) && !(mb & 0x08u);
        const uint32_t sus32 = __ballot_sync(JB_FULL, sus);
        if (!sus32) continue;
        const uint32_t slot = slot0 + (c0 + k) * 32 + lane;
        double lam_v[VM];
        double lam = 0.0;
#pragma unroll
        for (int v = 0; v < VM; ++v) lam_v[v] = 0.0;
        if (sus) lam = lambda_of(mb, slot, dim > 1 ? (int)JB_TI_LSG((uint32_t)(info8 >> (8 * k)) & 0xFFu) : 0, T, nj, beta, lam_v);
        const uint32_t ord = ord_base + running + __popc(sus32 & lane_lt);
        running += __popc(sus32);
        enqueue(sus && lam != 0.0, slot, lam_v, ord, g, slot0, n);
      }
    }
  }
  if (qn) jb_span_drain<VM, EV>(a, q, qn, absmask);
  if (lane == 0 && draws) atomicAdd(&a.counters[CNT_DRAWS], draws);
}
