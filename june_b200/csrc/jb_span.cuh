// Interaction kernel of a MIRRORED supergroup with ONE subgroup per group (companies: primary activity, static
// members, 1x1 contact matrix): Interaction.time_step_for_group (interaction.py:516-641) over a slot-ordered copy
// of the members' state bytes.  Included by jb_kernels.cuh after jb_groups.cuh.
//
// Everything a slot needs is read from arrays laid out in slot order -- the state byte (`marr`, kept current by
// the placement), the slot descriptor (`tp_info`), the member's person index (`tp_member`) and global id
// (`mgid`) -- so every load of the kernel is a coalesced stream; the only gathers left are the 8 bytes of an
// infector's transmission probability and of a susceptibility that is neither 0 nor 1.
//
// One WARP per work item, no block-level barrier.  A lane owns 16 consecutive slots: one 128-bit load of the
// state bytes (and of the descriptors), presence / infector / susceptible masks by byte-wise SWAR logic.
//   * a span of up to 16 tiles of small groups (<= 32 slots, never split by a tile): the two lanes of a tile
//     exchange their masks; every lane walks the groups that START in its half and hold both an infector and a
//     susceptible, sums the infectors' transmission probabilities in slot order (the reference's left-to-right
//     Python sum, interaction.py:463-467) and parks the group's Lambda per unit susceptibility in shared memory;
//   * 1 / 2 / 4 / 8 groups of 257-512 / 129-256 / 65-128 / 33-64 slots, each padded to its class size, so that a
//     group is a sub-warp of 32 / 16 / 8 / 4 lanes: size and sums by a butterfly inside the sub-warp (lane
//     partials in slot order);
//   * one group of more than 512 slots: first pass size and sums over its 512-slot chunks, second pass the
//     susceptibles.
// (jb_pack_span lays the supergroup out in exactly these work items, so every warp is full.)
// The susceptibles that draw are then taken tile by tile with lane == slot and compacted into the warp's draw
// queue; whenever 32 draws are waiting they are executed -- Philox + exp (the expensive part) always run on fully
// populated warps, and the kernel holds ONE out-of-line copy of them (it has to stay inside the instruction cache).
#pragma once

#define JB_SPAN_TILES 16                      // tiles per span of small groups (512 slots: 16 per lane)
// warps per CTA: the parking area is 4 KB per warp and variant, and static shared memory ends at 48 KB
#define JB_SPAN_WARPS_OF(VM) ((VM) == 1 ? 4 : 2)
#define JB_SPAN_QUEUE 64
#ifndef JB_SPAN_MINB
#define JB_SPAN_MINB 10  // resident CTAs per SM the single-variant kernel is compiled for (48 registers; measured: 8 -> 317 us, 10 -> 285 us, 12 -> 299 us at 56 M people)
#endif

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

// Bit 0 of each of the four bytes of x -> 4-bit mask.
__device__ __forceinline__ uint32_t jb_pack4(uint32_t x01) { return ((x01 * 0x00204081u) >> 21) & 0xFu; }

// Shared memory of the kernel: per warp the draw queue and the parking area of the small groups' row sums.
template <int VM, bool EV>
struct JbSpanSmem {
  static constexpr int JB_SPAN_WARPS = JB_SPAN_WARPS_OF(VM);
  uint32_t q_x[JB_SPAN_WARPS][JB_SPAN_QUEUE];      // slot of the susceptible (relative to the supergroup's first slot)
  uint32_t q_ord[JB_SPAN_WARPS][JB_SPAN_QUEUE];    // ordinal of the injected uniform (test mode)
  double q_row[JB_SPAN_WARPS][VM][JB_SPAN_QUEUE];  // Lambda per unit susceptibility
  // blame context of a queued draw (events only): group, its first slot, its slots
  uint32_t q_g[EV ? JB_SPAN_WARPS : 1][EV ? JB_SPAN_QUEUE : 1], q_s0[EV ? JB_SPAN_WARPS : 1][EV ? JB_SPAN_QUEUE : 1],
      q_len[EV ? JB_SPAN_WARPS : 1][EV ? JB_SPAN_QUEUE : 1];
  double row[JB_SPAN_WARPS][VM][JB_SPAN_TILES * 32];  // [slot of the span] valid at the first slot of a group
};
template <int VM, bool EV>
__device__ __forceinline__ JbSpanSmem<VM, EV>& jb_span_smem() {
  __shared__ JbSpanSmem<VM, EV> s;
  return s;
}

// Execute the first n (<= 32) queued draws of this warp.  Out of line on purpose -- see the header.
template <int VM, bool EV>
__device__ __noinline__ void jb_span_drain(const GroupArgs& a, uint32_t n, uint32_t absmask) {
  JbSpanSmem<VM, EV>& s = jb_span_smem<VM, EV>();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if ((uint32_t)lane < n) {
    const uint32_t x = s.q_x[wib][lane];
    const uint32_t code = a.marr[a.stream_base + x] & 0x03u;
    const bool need_p = VM > 1 || a.lambda != nullptr || code == JB_SC_READ;
    uint32_t p = need_p ? a.tp_member[x] : 0u;
    const int V = (VM > 1) ? a.V : 1;
    double lam_v[VM];
    double lam = 0.0;
#pragma unroll
    for (int v = 0; v < VM; ++v) {
      double t = 0.0;
      if (v < V) {
        const double sv = (VM == 1) ? (code == JB_SC_ONE ? 1.0 : code == JB_SC_ZERO ? 0.0 : a.susc[p]) : a.susc[(size_t)v * a.NP + p];
        t = s.q_row[wib][v][lane] * sv;
      }
      lam_v[v] = t;
      lam += t;
    }
    if (a.lambda) a.lambda[p] = lam;
    if (lam != 0.0) {  // lambda == 0 can never infect
      const int v = jb_draw_core<VM>(a, a.inject ? 0u : a.mgid[a.stream_base + x], lam_v, lam, s.q_ord[wib][lane]);
      if (v >= 0) {
        if (!need_p) p = a.tp_member[x];
        uint32_t infector = 0xFFFFFFFFu, g = 0;
        if (EV) {
          g = s.q_g[wib][lane];
          infector = jb_blame_span<VM>(a, jb_trans_read(a.trans, a.NP, a.sp), absmask, s.q_s0[wib][lane], s.q_len[wib][lane], x, g, v, p);
        }
        jb_emit(a, p, (uint32_t)v, g, infector);
      }
    }
  }
  __syncwarp();
}

template <int VM, bool EV>
__global__ void __launch_bounds__(32 * JB_SPAN_WARPS_OF(VM), VM == 1 && !EV ? JB_SPAN_MINB : 1) k_span(const __grid_constant__ GroupArgs a) {
  jb_stamp(3); jb_stamp(4);
  JbSpanSmem<VM, EV>& s = jb_span_smem<VM, EV>();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t item_idx = blockIdx.x * JB_SPAN_WARPS_OF(VM) + wib;
  if (item_idx >= a.n_span) return;
  const uint4 item = a.span_items[item_idx];
  const uint32_t kind = item.y >> 28;
  const double* const trans = jb_trans_read(a.trans, a.NP, a.sp);
  const uint32_t absmask = jb_mirror_absmask(a.sp->activity_mask, a.sp->flags);
  const uint32_t abs4 = absmask * 0x01010101u;
  const uint8_t* const marr = a.marr + a.stream_base;
  const int V = (VM > 1) ? a.V : 1;
  const double cm0 = a.cm[0], dt = a.sp->dt;
  const uint32_t lane_lt = (1u << lane) - 1u;
  // test modes (injected uniforms, recorded Lambda): every susceptible of a group that must timestep is queued, also
  // the immune ones (they consume a draw, interaction.py:610-641)
  const bool all_sus = a.inject != nullptr || a.lambda != nullptr;
  uint32_t qn = 0;     // warp-uniform queue length
  uint32_t draws = 0;  // per lane

  // The lanes with `want` append their draw; 32 waiting draws are executed.
  auto push = [&](uint32_t wm, uint32_t x, const double* row, uint32_t ord, uint32_t g, uint32_t s0, uint32_t len) {
    if ((wm >> lane) & 1u) {
      const uint32_t pos = qn + __popc(wm & lane_lt);
      jb_prefetch_l2(a.mgid + a.stream_base + x);  // the Philox counter word, read when the draw is executed
      s.q_x[wib][pos] = x;
      if (a.inject) s.q_ord[wib][pos] = ord;
      if (EV) { s.q_g[wib][pos] = g; s.q_s0[wib][pos] = s0; s.q_len[wib][pos] = len; }
#pragma unroll
      for (int v = 0; v < VM; ++v) s.q_row[wib][v][pos] = row[v];
    }
    qn += __popc(wm);
    __syncwarp();
    if (qn >= 32) {
      jb_span_drain<VM, EV>(a, 32u, absmask);
      const uint32_t rest = qn - 32;
      uint32_t mx = 0, mo = 0, mg = 0, ms = 0, ml_ = 0;
      double ml[VM];
      if ((uint32_t)lane < rest) {
        mx = s.q_x[wib][32 + lane]; mo = s.q_ord[wib][32 + lane];
        if (EV) { mg = s.q_g[wib][32 + lane]; ms = s.q_s0[wib][32 + lane]; ml_ = s.q_len[wib][32 + lane]; }
#pragma unroll
        for (int v = 0; v < VM; ++v) ml[v] = s.q_row[wib][v][32 + lane];
      }
      __syncwarp();
      if ((uint32_t)lane < rest) {
        s.q_x[wib][lane] = mx; s.q_ord[wib][lane] = mo;
        if (EV) { s.q_g[wib][lane] = mg; s.q_s0[wib][lane] = ms; s.q_len[wib][lane] = ml_; }
#pragma unroll
        for (int v = 0; v < VM; ++v) s.q_row[wib][v][lane] = ml[v];
      }
      __syncwarp();
      qn = rest;
    }
  };
  // 16 state bytes (four words) -> 16-bit masks: present, infector, susceptible, susceptible that draws
  auto masks16 = [&](const uint4& mv, uint32_t valid16, uint32_t& pres, uint32_t& inf, uint32_t& sus, uint32_t& el) {
    const uint32_t w[4] = {mv.x, mv.y, mv.z, mv.w};
    pres = 0; inf = 0; el = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t b = w[q], ab = b & abs4;
      const uint32_t absent01 = ((((ab & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | ab) >> 7) & 0x01010101u;
      const uint32_t pres01 = absent01 ^ 0x01010101u;
      const uint32_t inf01 = (b >> 3) & pres01;
      const uint32_t el01 = pres01 & ~inf01 & (b | (b >> 1));  // susceptibility code != 0
      pres |= jb_pack4(pres01) << (4 * q);
      inf |= jb_pack4(inf01) << (4 * q);
      el |= jb_pack4(el01) << (4 * q);
    }
    pres &= valid16; inf &= valid16;
    sus = pres & ~inf;
    el = all_sus ? sus : (el & valid16);
  };
  // transmission probability of the infectors in `m` (bits relative to slot `base`), summed in slot order
  auto sum_trans = [&](uint32_t m, uint32_t base, double (&T)[VM]) {
    for (; m; m &= m - 1u) {
      const uint32_t sl = base + (uint32_t)__ffs(m) - 1u;
      const double t = trans[a.tp_member[sl]];
      if (VM == 1) {
        T[0] += t;
      } else {
        const uint32_t vk = jb_mb_var(marr[sl]);
#pragma unroll
        for (int v = 0; v < VM; ++v)
          if ((uint32_t)v == vk) T[v] += t;
      }
    }
  };

  const uint32_t slot0 = item.x;
  if (kind == JB_SPAN_KIND_SMALL) {
    // ---- a span of tiles of small groups ---------------------------------------------------------------------
    const uint32_t nt = item.y & 0xFFu;
    const bool ok = (uint32_t)(lane >> 1) < nt;
    const uint32_t my0 = slot0 + lane * 16;  // the lane's first slot
    uint4 mv = make_uint4(0u, 0u, 0u, 0u), iv = make_uint4(0u, 0u, 0u, 0u);
    uint32_t tg0 = 0;
    if (ok) {
      mv = *reinterpret_cast<const uint4*>(marr + my0);
      iv = *reinterpret_cast<const uint4*>(a.tp_info + my0);
      tg0 = a.tile_group0[(slot0 >> 5) + (lane >> 1)];
    }
    uint32_t valid16 = 0, first16 = 0;
    {
      const uint32_t w[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t valid01 = (w[q] >> 3) & 0x01010101u;
        valid16 |= jb_pack4(valid01) << (4 * q);
        first16 |= jb_pack4((w[q] >> 4) & valid01) << (4 * q);
      }
    }
    uint32_t pres16, inf16, sus16, el16;
    masks16(mv, valid16, pres16, inf16, sus16, el16);
    // a tile = the two neighbouring lanes: both get the tile's 32-bit masks
    const uint32_t sh = (lane & 1) * 16, osh = 16 - sh;
    const uint32_t o1 = __shfl_xor_sync(JB_FULL, pres16 | (inf16 << 16), 1);
    const uint32_t o2 = __shfl_xor_sync(JB_FULL, el16 | (first16 << 16), 1);
    const uint32_t o3 = __shfl_xor_sync(JB_FULL, valid16, 1);
    const uint32_t pres32 = (pres16 << sh) | ((o1 & 0xFFFFu) << osh), inf32 = (inf16 << sh) | ((o1 >> 16) << osh);
    const uint32_t el32 = (el16 << sh) | ((o2 & 0xFFFFu) << osh), first32 = (first16 << sh) | ((o2 >> 16) << osh);
    const uint32_t n_valid = __popc((valid16 << sh) | (o3 << osh));  // valid slots are a prefix of the tile
    const uint32_t sus32 = pres32 & ~inf32;
    const uint32_t tl = (uint32_t)(lane >> 1) * 32u;  // the tile's first slot inside the span
    uint32_t rem = inf32;  // infectors whose group has not been looked at yet
    uint32_t umask = 0;    // susceptibles that draw, of the groups this lane handled
    while (rem) {
      // next group that starts in this lane's half of the tile and holds an infector and a susceptible
      const uint32_t src = __ffs(rem) - 1;
      const uint32_t le = 0xFFFFFFFFu >> (31 - src);
      const uint32_t below = first32 & le, above = first32 & ~le;
      const uint32_t start = 31 - __clz(below | 1u);
      const uint32_t end = above ? (uint32_t)__ffs(above) - 1u : n_valid;
      const uint32_t gmask = (end >= 32u ? 0xFFFFFFFFu : ((1u << end) - 1u)) & ~((1u << start) - 1u);
      rem &= ~gmask;
      if ((start >> 4) != (uint32_t)(lane & 1) || !(gmask & sus32)) continue;  // interactive.py:136
      const uint32_t g = a.span_group[tg0 + __popc(below) - 1u];
      const uint32_t nsus = __popc(gmask & sus32);
      if (a.mode == 1) { a.draw_cnt[g] = nsus; continue; }  // only count the draws per group (injected-uniform ordinals)
      draws += nsus;
      double T[VM];
#pragma unroll
      for (int v = 0; v < VM; ++v) T[v] = 0.0;
      sum_trans(gmask & inf32, slot0 + tl, T);
      const uint32_t gi = a.grp_info[g];
      const double beta = a.beta[JB_GINFO_BETA(gi) * a.R + JB_GINFO_REGION(gi)] * a.beta_scale[JB_GINFO_SCALE(gi)];
      const int nn = max(1, (int)__popc(gmask & pres32) - 1);  // interaction.py:469-471
#pragma unroll
      for (int v = 0; v < VM; ++v) s.row[wib][v][tl + start] = (v < V && T[v] != 0.0) ? cm0 * T[v] / (double)nn * beta * dt : 0.0;
      umask |= gmask & el32;
    }
    umask |= __shfl_xor_sync(JB_FULL, umask, 1);
    __syncwarp();
    if (a.mode != 1) {
      // tile by tile, lane == slot: the susceptibles that draw find their group's row sum and join the queue
#pragma unroll 1
      for (uint32_t t = 0; t < nt; ++t) {
        const uint32_t U = __shfl_sync(JB_FULL, umask, 2 * t);
        if (!U) continue;
        const uint32_t f = __shfl_sync(JB_FULL, first32, 2 * t);
        const uint32_t start = 31 - __clz((f & (0xFFFFFFFFu >> (31 - lane))) | 1u);
        double row[VM];
#pragma unroll
        for (int v = 0; v < VM; ++v) row[v] = ((U >> lane) & 1u) ? s.row[wib][v][t * 32 + start] : 0.0;
        uint32_t ord = 0, g = 0, s0 = 0, len = 0;
        if (a.inject || EV) {
          const uint32_t tg0_t = __shfl_sync(JB_FULL, tg0, 2 * t);
          g = a.span_group[tg0_t + __popc(f & (0xFFFFFFFFu >> (31 - lane))) - 1u];
          if (a.inject) {  // ordinal = the group's susceptibles before this slot (slot order)
            const uint32_t su = __shfl_sync(JB_FULL, sus32, 2 * t);
            ord = (((U >> lane) & 1u) ? a.draw_base[g] : 0u) + __popc(su & lane_lt & ~((1u << start) - 1u));
          }
          if (EV) {
            const uint32_t nv = __shfl_sync(JB_FULL, n_valid, 2 * t);
            const uint32_t above = f & ~(0xFFFFFFFFu >> (31 - lane));
            s0 = slot0 + t * 32 + start;
            len = (above ? (uint32_t)__ffs(above) - 1u : nv) - start;
          }
        }
        push(U, slot0 + t * 32 + lane, row, ord, g, s0, len);
      }
    }
  } else if (kind == JB_SPAN_KIND_CLASS) {
    // ---- groups of one size class: a group = a sub-warp of W lanes (16 slots per lane) --------------------------
    const uint32_t W = (item.y >> 8) & 0xFFu, cnt = item.y & 0xFFu;
    const uint32_t j = (uint32_t)lane / W, li = (uint32_t)lane % W;
    const bool have = j < cnt;
    const uint32_t g = have ? a.span_group[item.z + j] : 0u;
    const uint32_t n = have ? a.grp_off[g + 1] - a.grp_off[g] : 0u;
    const uint32_t idx0 = li * 16u, my0 = slot0 + lane * 16u;
    const uint32_t valid16 = idx0 + 16u <= n ? 0xFFFFu : idx0 < n ? (1u << (n - idx0)) - 1u : 0u;
    const uint4 mv = valid16 ? *reinterpret_cast<const uint4*>(marr + my0) : make_uint4(0u, 0u, 0u, 0u);
    uint32_t pres16, inf16, sus16, el16;
    masks16(mv, valid16, pres16, inf16, sus16, el16);
    double T[VM];
#pragma unroll
    for (int v = 0; v < VM; ++v) T[v] = 0.0;
    sum_trans(inf16, my0, T);
    // size | susceptibles | infectors of the group (10 bits each), summed transmission probabilities: sub-warp butterfly
    uint32_t cnt3 = (uint32_t)__popc(pres16) | ((uint32_t)__popc(sus16) << 10) | ((uint32_t)__popc(inf16) << 20);
    for (uint32_t off = 1; off < W; off <<= 1) {
      cnt3 += __shfl_xor_sync(JB_FULL, cnt3, off);
#pragma unroll
      for (int v = 0; v < VM; ++v) T[v] += __shfl_xor_sync(JB_FULL, T[v], off);
    }
    const uint32_t npres = cnt3 & 0x3FFu, nsus = (cnt3 >> 10) & 0x3FFu, ninf = cnt3 >> 20;
    const bool must = nsus > 0 && ninf > 0;  // interactive.py:136
    if (a.mode == 1) {
      if (have && li == 0) a.draw_cnt[g] = must ? nsus : 0u;
      return;
    }
    double row[VM];
#pragma unroll
    for (int v = 0; v < VM; ++v) row[v] = 0.0;
    if (must) {
      if (li == 0) draws += nsus;
      const uint32_t gi = a.grp_info[g];
      const double beta = a.beta[JB_GINFO_BETA(gi) * a.R + JB_GINFO_REGION(gi)] * a.beta_scale[JB_GINFO_SCALE(gi)];
      const int nn = max(1, (int)npres - 1);
#pragma unroll
      for (int v = 0; v < VM; ++v) row[v] = (v < V && T[v] != 0.0) ? cm0 * T[v] / (double)nn * beta * dt : 0.0;
    } else {
      el16 = 0;
    }
    uint32_t ordl = 0;  // injected uniforms: ordinal of the first susceptible of this lane's slots
    if (a.inject) {
      uint32_t incl = __popc(sus16);
      for (uint32_t off = 1; off < W; off <<= 1) {
        const uint32_t up = __shfl_up_sync(JB_FULL, incl, off);
        if (li >= off) incl += up;
      }
      ordl = (must ? a.draw_base[g] : 0u) + incl - __popc(sus16);
    }
    // tile by tile, lane == slot: tile t belongs to the group of lane 2t
#pragma unroll 1
    for (uint32_t t = 0; t < 16u; ++t) {
      const uint32_t U = __shfl_sync(JB_FULL, el16, 2 * t) | (__shfl_sync(JB_FULL, el16, 2 * t + 1) << 16);
      if (!U) continue;
      double row_t[VM];
#pragma unroll
      for (int v = 0; v < VM; ++v) row_t[v] = __shfl_sync(JB_FULL, row[v], 2 * t);
      uint32_t ord = 0, g_t = 0, n_t = 0;
      if (a.inject) {
        const uint32_t su = __shfl_sync(JB_FULL, sus16, 2 * t) | (__shfl_sync(JB_FULL, sus16, 2 * t + 1) << 16);
        ord = __shfl_sync(JB_FULL, ordl, 2 * t) + __popc(su & lane_lt);
      }
      if (EV) { g_t = __shfl_sync(JB_FULL, g, 2 * t); n_t = __shfl_sync(JB_FULL, n, 2 * t); }
      push(U, slot0 + t * 32 + lane, row_t, ord, g_t, slot0 + ((2 * t) / W) * W * 16u, n_t);
    }
  } else {
    // ---- one group of more than 512 slots ------------------------------------------------------------------------
    const uint32_t g = a.span_group[item.z];
    const uint32_t n = a.grp_off[g + 1] - a.grp_off[g];
    const uint32_t nchunk = (n + 511u) >> 9;
    uint4 mv = make_uint4(0u, 0u, 0u, 0u);
    uint32_t pres16 = 0, inf16 = 0, sus16 = 0, el16 = 0;
    auto load_chunk = [&](uint32_t c) {
      const uint32_t idx0 = c * 512u + lane * 16u;
      const uint32_t valid16 = idx0 + 16u <= n ? 0xFFFFu : idx0 < n ? (1u << (n - idx0)) - 1u : 0u;
      mv = valid16 ? *reinterpret_cast<const uint4*>(marr + slot0 + idx0) : make_uint4(0u, 0u, 0u, 0u);
      masks16(mv, valid16, pres16, inf16, sus16, el16);
    };
    // pass 1: size, susceptibles, summed transmission probabilities (lane partials in slot order)
    uint32_t npres = 0, nsus = 0, ninf = 0;
    double T[VM];
#pragma unroll
    for (int v = 0; v < VM; ++v) T[v] = 0.0;
    for (uint32_t c = 0; c < nchunk; ++c) {
      load_chunk(c);
      npres += __popc(pres16); nsus += __popc(sus16); ninf += __popc(inf16);
      sum_trans(inf16, slot0 + c * 512u + lane * 16u, T);
    }
    npres = __reduce_add_sync(JB_FULL, npres); nsus = __reduce_add_sync(JB_FULL, nsus); ninf = __reduce_add_sync(JB_FULL, ninf);
    const bool must = nsus > 0 && ninf > 0;  // interactive.py:136
    if (a.mode == 1) {
      if (lane == 0) a.draw_cnt[g] = must ? nsus : 0u;
      return;
    }
    if (!must) return;
    if (lane == 0) draws += nsus;
#pragma unroll
    for (int v = 0; v < VM; ++v)
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) T[v] += __shfl_xor_sync(JB_FULL, T[v], off);  // fixed order: same bits in every lane
    const uint32_t gi = a.grp_info[g];
    const double beta = a.beta[JB_GINFO_BETA(gi) * a.R + JB_GINFO_REGION(gi)] * a.beta_scale[JB_GINFO_SCALE(gi)];
    const int nn = max(1, (int)npres - 1);
    double row[VM];
#pragma unroll
    for (int v = 0; v < VM; ++v) row[v] = (v < V && T[v] != 0.0) ? cm0 * T[v] / (double)nn * beta * dt : 0.0;
    // pass 2: tile by tile, lane == slot (a group of one chunk is still in registers)
    uint32_t running = a.inject ? a.draw_base[g] : 0u;  // ordinal of the next tile's first susceptible
    for (uint32_t c = 0; c < nchunk; ++c) {
      if (nchunk > 1) load_chunk(c);
      const uint32_t ntl = min(16u, (n - c * 512u + 31u) >> 5);
#pragma unroll 1
      for (uint32_t t = 0; t < ntl; ++t) {
        const uint32_t U = __shfl_sync(JB_FULL, el16, 2 * t) | (__shfl_sync(JB_FULL, el16, 2 * t + 1) << 16);
        uint32_t ord = 0;
        if (a.inject) {
          const uint32_t su = __shfl_sync(JB_FULL, sus16, 2 * t) | (__shfl_sync(JB_FULL, sus16, 2 * t + 1) << 16);
          ord = running + __popc(su & lane_lt);
          running += __popc(su);
        }
        if (!U) continue;
        push(U, slot0 + c * 512u + t * 32u + lane, row, ord, g, slot0, n);
      }
    }
  }
  if (qn) jb_span_drain<VM, EV>(a, qn, absmask);
  draws = __reduce_add_sync(JB_FULL, draws);
  if (lane == 0 && draws) atomicAdd(&a.counters[CNT_DRAWS], draws);
}
