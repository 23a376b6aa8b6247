// Interaction kernels (Interaction.time_step_for_group, interaction.py:516-641, with
// InteractiveGroup.__init__, interactive.py:37-137).  Included by jb_kernels.cuh.
//
//   k_groups       one warp or one CTA per group: schools, hospitals (dynamic members), any group
//                  with more than 32 registered slots or a contact matrix larger than 4x4
//   k_tiles        groups of <= 32 slots packed into 32-slot tiles: streaming detection of the
//                  groups that must timestep + their processing, fused in one kernel through a
//                  shared-memory work list
//
// Per present member the kernels read ONE byte (`phot`: activity, infected flag, susceptibility
// code, variant); 8 more bytes only for infectors (transmission probability) and for
// susceptibles whose susceptibility is neither 0 nor 1.
#pragma once

// Contact matrix element for (susceptible subgroup i, infector subgroup j).
__device__ __forceinline__ double jb_cm_elem(const GroupArgs& a, int i, int j, const uint8_t* years, int n_years) {
  if (a.kind == JB_KIND_MATRIX) return a.cm[i * a.dim + j];
  // school.py:462-485 on the 32x32 table built by school.py:424-460
  const double* c = a.cm;
  if (i == j) return i != 0 ? c[1 * JB_SCHOOL_DIM + 1] : c[0];
  if (i == 0) return c[0 * JB_SCHOOL_DIM + 1] / (double)n_years;
  if (j == 0) return c[1 * JB_SCHOOL_DIM + 0] / (double)n_years;
  int yi = years[i - 1] + 1, yj = years[j - 1] + 1;
  double x = c[yi * JB_SCHOOL_DIM + yj];
  return yi == yj ? x / 4.0 : x;
}

// Susceptibility of person p to variant v given its hot byte.
template <int VM>
__device__ __forceinline__ double jb_susc(const GroupArgs& a, uint32_t p, uint32_t hot, int v) {
  if (VM == 1) {
    const uint32_t code = JB_HOT_CODE(hot);
    return code == JB_SC_ONE ? 1.0 : code == JB_SC_ZERO ? 0.0 : a.susc[p];
  }
  return a.susc[(size_t)v * a.NP + p];
}

// Inclusive segmented scan over a warp for keys that are contiguous (sorted) in lane order.
template <typename Tv>
__device__ __forceinline__ Tv jb_seg_scan(Tv val, uint32_t key, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    Tv up = __shfl_up_sync(JB_FULL, val, off);
    uint32_t kup = __shfl_up_sync(JB_FULL, key, off);
    if (lane >= off && kup == key) val += up;
  }
  return val;
}

// A member as two registers: person index and {lsg (16 bits) | hot byte << 16 | present << 30 |
// infected << 31}.
struct JbMember {
  uint32_t p, k;
  __device__ __forceinline__ uint32_t lsg() const { return k & 0xFFFFu; }
  __device__ __forceinline__ uint32_t hot() const { return (k >> 16) & 0xFFu; }
  __device__ __forceinline__ bool present() const { return (k >> 30) & 1u; }
  __device__ __forceinline__ bool infector() const { return (k >> 30) == 3u; }   // present and infected
  __device__ __forceinline__ bool susceptible() const { return (k >> 30) == 1u; }  // present, not infected
};

// What the member loads of one group need besides the launch arguments.
struct JbGrp {
  const double* trans;
  const uint32_t* dyn_s;    // the group's sorted dynamic members (shared memory)
  uint32_t b, ns, ntot;     // first registration, static members, static + dynamic members
  uint32_t slot0, absmask;  // mirrored supergroup: the group's first slot in `marr`, absence bits of this step
};

// Member `idx` of group (static registered slots first, then the sorted dynamic members, which
// the CTA has staged in shared memory as (lsg << 28) | person, sorted).
// In a mirrored supergroup the member's byte is read from the slot-ordered array: coalesced, and in the same
// round trip as the member index instead of behind it.
__device__ __forceinline__ JbMember jb_load_member(const GroupArgs& a, const JbGrp& c, uint32_t idx) {
  JbMember mbr;
  mbr.p = 0; mbr.k = 0xFFFFu;
  if (idx >= c.ntot) return mbr;
  if (idx < c.ns) {
    const uint32_t m = c.b + idx;
    const uint32_t info = a.reg_info[m];
    mbr.p = a.reg_member[m];
    const uint32_t hot = a.marr ? jb_mirror_to_hot(a.marr[c.slot0 + idx], c.absmask, JB_INFO_ACT(info)) : a.phot[mbr.p];
    const uint32_t pres = (hot & JB_PS_ACT_MASK) == JB_INFO_ACT(info) ? 1u : 0u;
    mbr.k = JB_INFO_LSG(info) | (hot << 16) | (pres << 30) | ((hot & JB_PS_INFECTED) ? 1u << 31 : 0u);
  } else {
    const uint32_t k = c.dyn_s[idx - c.ns];  // (lsg << 28) | caller's person index
    mbr.p = a.ext2int[k & 0x0FFFFFFFu];
    // (few, scattered members: their value and id start their trips together with the hot byte's; for the
    // static members of the big company lists the extra sectors cost more than the latency they hide)
    jb_prefetch_l2(c.trans + mbr.p);
    if (mbr.p < a.N) jb_prefetch_l2(a.int2ext + mbr.p);
    const uint32_t hot = a.phot[mbr.p];
    mbr.k = (k >> 28) | (hot << 16) | (1u << 30) | ((hot & JB_PS_INFECTED) ? 1u << 31 : 0u);
  }
  return mbr;
}

// ---- group kernel: WPG warps per group --------------------------------------------------------
// WPG = 1: one warp per group, JB_CTA_WARPS groups per CTA (groups of <= 128 members: the whole
// group sits in the warp's registers between the two passes).  WPG = JB_CTA_WARPS: one CTA per
// group (large groups, and groups with dynamic members whose bin the CTA sorts).
#define JB_CTA 128
#define JB_CTA_WARPS (JB_CTA / 32)
#define JB_UB 4  // 32-member batches in flight per warp: the member -> byte -> value chain is latency bound

// Shared memory per group: n[L] u32 | sus_static[L] | sus_dyn[L] | totals[2] | pad | T[WPG][V][L]
// f64 (one copy per warp: fixed summation order) | row[V][L] f64 | dynamic member staging
__host__ __device__ inline size_t jb_grp_smem_bytes(int L, int V, uint32_t bin_cap, int wpg) {
  size_t ints = ((size_t)3 * L + 2 + 1) & ~(size_t)1;
  size_t b = ints * 4 + (size_t)(wpg + 1) * V * L * 8 + (size_t)bin_cap * 4;
  return (b + 7) & ~(size_t)7;
}

template <int WPG>
__device__ __forceinline__ void jb_grp_sync() {
  if (WPG == 1) __syncwarp(); else __syncthreads();
}

// Rare path (a draw succeeded and events are recorded), kept out of line so that it does not
// cost the hot loop registers.  _blame_subgroup (interaction.py:643-645): infector subgroup j with
// probability ~ its term of the susceptible's row; _blame_individuals (:647-662): infector with
// probability ~ transmission probability inside that subgroup, in member order.
template <int VM>
__device__ __noinline__ uint32_t jb_blame_group(const GroupArgs& a, const JbGrp& gc, uint32_t p, int i, int v, int Lg, const uint8_t* years,
                                                int n_years, const uint32_t* n_l, const double* T, double row,
                                                double beta, double dt) {
  const double* trans = gc.trans;
  const uint32_t ntot = gc.ntot;
  double u_sub, u_ind;
  jb_blame_uniforms(a, p, u_sub, u_ind);
  const double target = u_sub * row;
  double c = 0.0;
  int jb = -1, jlast = 0;
  for (int j = 0; j < Lg && jb < 0; ++j) {
    const double t = T[j];
    if (t == 0.0) continue;
    const int nj = (int)n_l[j];
    const int nn = (i == j) ? max(1, nj - 1) : nj;
    c += jb_cm_elem(a, i, j, years, n_years) * t / (double)nn * beta * dt;
    jlast = j;
    if (c > target) jb = j;
  }
  if (jb < 0) jb = jlast;
  const double target2 = u_ind * T[jb];
  double c2 = 0.0;
  uint32_t infector = 0xFFFFFFFFu;
  for (uint32_t idx = 0; idx < ntot; ++idx) {
    const JbMember o = jb_load_member(a, gc, idx);
    if (!o.infector() || (int)o.lsg() != jb || (VM > 1 && (int)JB_HOT_VAR(o.hot()) != v)) continue;
    c2 += trans[o.p];
    infector = o.p;
    if (c2 > target2) break;
  }
  return infector;
}

// One group (entry `gi_` of the launch's list); called by the threads that share it (a warp, or the CTA).
template <int VM, int WPG, bool EV>
__device__ __forceinline__ void jb_group_one(const GroupArgs& a, const uint32_t gi_) {
  const double* const trans = jb_trans_read(a.trans, a.NP, a.sp);
  extern __shared__ __align__(8) unsigned char smem_raw[];
  constexpr int GT = 32 * WPG;                 // threads per group
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int wl = wib % WPG, gs = wib / WPG, tg = threadIdx.x - gs * GT;
  const uint32_t g = a.group_list ? a.group_list[gi_] : a.g0 + gi_;
  // a household that is being visited belongs to the visited-list launch (which sets dyn_by_list)
  if (a.vslot && !a.dyn_by_list && a.vslot[g - a.g0] != JB_VIS_NONE) return;
  const uint32_t bi = a.dyn_by_list ? gi_ : g - a.g0;  // which bin holds the group's dynamic members
  const int L = a.L, V = (VM > 1) ? a.V : 1;
  const uint32_t bin_cap = a.dyn_cnt ? a.bin_cap : 0u;
  unsigned char* base = smem_raw + (size_t)gs * jb_grp_smem_bytes(L, V, bin_cap, WPG);
  uint32_t* n_l = (uint32_t*)base;
  uint32_t* ss_l = n_l + L;   // susceptibles among static members, per subgroup
  uint32_t* sd_l = ss_l + L;  // susceptibles among dynamic members
  uint32_t* tot = sd_l + L;   // [0] susceptibles, [1] infectors
  double* T_w = (double*)(base + ((((size_t)3 * L + 2 + 1) & ~(size_t)1) * 4));
  double* row_l = T_w + (size_t)WPG * V * L;
  uint32_t* dyn_s = (uint32_t*)(row_l + (size_t)V * L);

  // registered slots can only hold someone when their feeding activity takes place in this step
  // (hospital workers on a residence-only step: the group is just its patients)
  // everything that only depends on the group index starts its trip now, instead of one DRAM latency after
  // the other along the kernel (offsets, bin fill and bin, group word)
  if (tg == 0) {
    jb_prefetch_l2(a.grp_info + g);
    if (a.dyn_cnt) {
      jb_prefetch_l2(a.dyn_cnt + a.dyn_bin_base + bi);
      jb_prefetch_l2(a.dyn_bin + (size_t)bi * bin_cap);
    }
  }
  const bool any_static = (a.static_acts & a.sp->activity_mask) != 0u;
  const uint32_t b = a.grp_off[g], ns = any_static ? a.grp_off[g + 1] - b : 0u;
  uint32_t nd = 0;
  if (a.dyn_cnt) {
    // stage this group's bin and sort it: members arrive in arbitrary (atomic) order, the
    // reference order is (subgroup, person) -- bitonic sort in shared memory
    nd = min(a.dyn_cnt[a.dyn_bin_base + bi], bin_cap);
    if (nd) {
      uint32_t npad = 32;
      while (npad < nd) npad <<= 1;
      const uint32_t* src = a.dyn_bin + (size_t)bi * bin_cap;  // dyn_bin = the supergroup's first bin
      for (uint32_t i = tg; i < npad; i += GT) dyn_s[i] = i < nd ? src[i] : 0xFFFFFFFFu;
      jb_grp_sync<WPG>();
      for (uint32_t k = 2; k <= npad; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          for (uint32_t i = tg; i < npad; i += GT) {
            const uint32_t ixj = i ^ j;
            if (ixj > i) {
              const uint32_t x = dyn_s[i], y = dyn_s[ixj];
              const bool asc = (i & k) == 0;
              if ((x > y) == asc) { dyn_s[i] = y; dyn_s[ixj] = x; }
            }
          }
          jb_grp_sync<WPG>();
        }
    }
  }
  const uint32_t ntot = ns + nd;
  if (ntot == 0) {
    if (a.mode == 1 && tg == 0) a.draw_cnt[g] = 0;
    return;
  }
  JbGrp c;
  c.trans = trans; c.dyn_s = dyn_s; c.b = b; c.ns = ns; c.ntot = ntot;
  c.slot0 = a.marr ? a.stream_base + a.list_slot[gi_] : 0u;
  c.absmask = a.marr ? jb_mirror_absmask(a.sp->activity_mask, a.sp->flags) : 0u;
  // number of local subgroups of this group
  int Lg = a.dim;
  const uint8_t* years = nullptr;
  int n_years = 0;
  if (a.kind == JB_KIND_SCHOOL) {
    const uint32_t aux = a.grp_aux[g];
    n_years = aux & 0xFFu;
    years = a.school_years + (aux >> 8);
    Lg = n_years + 1;
  }
  for (int i = tg; i < Lg; i += GT) {
    n_l[i] = 0; ss_l[i] = 0; sd_l[i] = 0;
    for (int w = 0; w < WPG; ++w)
      for (int v = 0; v < V; ++v) T_w[((size_t)w * V + v) * L + i] = 0.0;
  }
  if (tg < 2) tot[tg] = 0;
  jb_grp_sync<WPG>();

  // ---- pass 1: subgroup sizes and summed transmission probabilities ----------------------
  // Static and dynamic members are walked in separate batches so that the subgroup keys stay
  // sorted inside every 32-wide batch (needed by the segmented scan).  Warp w takes the batches
  // w, w + WPG, ... and accumulates into its own copy of T, so the summation order is fixed.
  // A group without dynamic members that fits in one round of batches stays in registers.
  const bool resident = nd == 0 && ns <= (uint32_t)(GT * JB_UB);
  JbMember keep[JB_UB];
  double* T_mine = T_w + (size_t)wl * V * L;
  uint32_t nsus = 0, ninf = 0;
  for (int phase = 0; phase < 2; ++phase) {
    const uint32_t lo = phase ? ns : 0u, hi = phase ? ntot : ns;
    for (uint32_t it0 = lo + wl * 32 * JB_UB; it0 < hi; it0 += GT * JB_UB) {
      JbMember mbs[JB_UB];
      double tvs[JB_UB];
#pragma unroll
      for (int u = 0; u < JB_UB; ++u) {
        const uint32_t idx = it0 + 32 * u + lane;
        mbs[u] = jb_load_member(a, c, idx < hi ? idx : ntot);
      }
#pragma unroll
      for (int u = 0; u < JB_UB; ++u) tvs[u] = mbs[u].infector() ? trans[mbs[u].p] : 0.0;
      if (resident) {
#pragma unroll
        for (int u = 0; u < JB_UB; ++u) keep[u] = mbs[u];
      }
#pragma unroll
      for (int u = 0; u < JB_UB; ++u) {
        if (it0 + 32 * u >= hi) break;
        const JbMember mb = mbs[u];
        const uint32_t key = mb.lsg();
        const uint32_t sus = mb.susceptible() ? 1u : 0u;
        uint32_t packed = (mb.present() ? 1u : 0u) | (sus << 16);
        packed = jb_seg_scan<uint32_t>(packed, key, lane);
        const uint32_t knext = __shfl_down_sync(JB_FULL, key, 1);
        const bool tail = (lane == 31 || knext != key) && key < (uint32_t)Lg;
        if (tail && packed) {
          atomicAdd(&n_l[key], packed & 0xFFFFu);
          if (packed >> 16) atomicAdd(phase == 0 ? &ss_l[key] : &sd_l[key], packed >> 16);
        }
        nsus += __popc(__ballot_sync(JB_FULL, sus));
        const bool inf = mb.infector();
        const uint32_t infm = __ballot_sync(JB_FULL, inf);
        ninf += __popc(infm);
        if (infm) {
          const int pv = (VM > 1 && inf) ? (int)JB_HOT_VAR(mb.hot()) : 0;
          for (int v = 0; v < V; ++v) {
            double tv = (pv == v) ? tvs[u] : 0.0;
            tv = jb_seg_scan<double>(tv, key, lane);
            if (tail) T_mine[v * L + key] += tv;
          }
          __syncwarp();
        }
      }
    }
  }
  if (WPG > 1) {
    if (lane == 0) {
      if (nsus) atomicAdd(&tot[0], nsus);
      if (ninf) atomicAdd(&tot[1], ninf);
    }
    __syncthreads();
    nsus = tot[0]; ninf = tot[1];
  } else {
    __syncwarp();
  }
  const bool must = nsus > 0 && ninf > 0;  // interactive.py:136
  if (a.mode == 1) {
    if (tg == 0) a.draw_cnt[g] = must ? nsus : 0;
    return;
  }
  if (!must) return;

  // ---- infector tensor row sums (interaction.py:451-478, :619) ---------------------------
  if (WPG > 1) {
    for (int i = tg; i < Lg * V; i += GT) {  // fold the per-warp copies, warp order
      const int v = i / Lg, j = i - v * Lg;
      double s = 0.0;
      for (int w = 0; w < WPG; ++w) s += T_w[((size_t)w * V + v) * L + j];
      T_w[(size_t)v * L + j] = s;
    }
    __syncthreads();
  }
  const uint32_t gi = a.grp_info[g];
  // (a household that is being visited takes the "household_visits" rule, household.py:279-289)
  const uint32_t rule = a.beta_rule ? (uint32_t)a.beta_rule - 1u : JB_GINFO_BETA(gi);
  const double beta = a.beta[rule * a.R + JB_GINFO_REGION(gi)] * a.beta_scale[JB_GINFO_SCALE(gi)];
  const double dt = a.sp->dt;
  for (int i = tg; i < Lg; i += GT) {
    for (int v = 0; v < V; ++v) {
      double s = 0.0;
      for (int j = 0; j < Lg; ++j) {
        const double t = T_w[(size_t)v * L + j];
        if (t != 0.0) {
          const int nj = (int)n_l[j];
          const int nn = (i == j) ? max(1, nj - 1) : nj;
          s += jb_cm_elem(a, i, j, years, n_years) * t / (double)nn * beta * dt;
        }
      }
      row_l[v * L + i] = s;
    }
  }
  // injected-uniform ordinals: susceptibles are numbered subgroup by subgroup, static members
  // before dynamic ones (interactive.py:86-131 dict insertion order)
  uint32_t ord_base = 0;
  if (a.inject) {
    ord_base = a.draw_base[g];
    if (tg == 0) {
      uint32_t run = 0;
      for (int i = 0; i < Lg; ++i) {
        const uint32_t s_ = ss_l[i], d_ = sd_l[i];
        ss_l[i] = run; run += s_;
        sd_l[i] = run; run += d_;
      }
    }
  }
  jb_grp_sync<WPG>();

  // ---- pass 2: one Bernoulli draw per susceptible ----------------------------------------
  // With injected uniforms the ordinal of a draw depends on all earlier members: warp 0 walks the
  // group alone, in order (test mode).  Otherwise the warps share the batches.
  const int n_walk = a.inject ? 1 : WPG;
  if (wl < n_walk) {
    for (int phase = 0; phase < 2; ++phase) {
      const uint32_t lo = phase ? ns : 0u, hi = phase ? ntot : ns;
      uint32_t* run = phase ? sd_l : ss_l;
      for (uint32_t it0 = lo + wl * 32 * JB_UB; it0 < hi; it0 += n_walk * 32 * JB_UB) {
        JbMember mbs[JB_UB];
        double s0[JB_UB];
        if (resident && n_walk == WPG) {
#pragma unroll
          for (int u = 0; u < JB_UB; ++u) mbs[u] = keep[u];
        } else {
#pragma unroll
          for (int u = 0; u < JB_UB; ++u) {
            const uint32_t idx = it0 + 32 * u + lane;
            mbs[u] = jb_load_member(a, c, idx < hi ? idx : ntot);
          }
        }
#pragma unroll
        for (int u = 0; u < JB_UB; ++u)
          s0[u] = (mbs[u].susceptible() && mbs[u].lsg() < (uint32_t)Lg) ? jb_susc<VM>(a, mbs[u].p, mbs[u].hot(), 0) : 0.0;
#pragma unroll
        for (int u = 0; u < JB_UB; ++u) {
          if (it0 + 32 * u >= hi) break;
          const JbMember mb = mbs[u];
          const bool sus = mb.susceptible() && mb.lsg() < (uint32_t)Lg;
          const uint32_t midx = it0 + 32 * u + lane;  // member index inside the group
          uint32_t ordinal = 0;
          if (a.inject) {
            const uint32_t key = mb.lsg();
            const uint32_t incl = jb_seg_scan<uint32_t>(sus ? 1u : 0u, key, lane);
            const uint32_t knext = __shfl_down_sync(JB_FULL, key, 1);
            const bool tail = (lane == 31 || knext != key) && key < (uint32_t)Lg;
            if (sus) ordinal = ord_base + run[key] + incl - 1;
            __syncwarp();
            if (tail) run[key] += incl;
            __syncwarp();
          }
          if (sus) {
            double lam_v[VM];
            double lam = 0.0;
#pragma unroll
            for (int v = 0; v < VM; ++v) {
              double r = 0.0, sv = 0.0;
              if (v < V) { r = row_l[v * L + mb.lsg()]; sv = (v == 0) ? s0[u] : a.susc[(size_t)v * a.NP + mb.p]; }
              lam_v[v] = r * sv;
              lam += lam_v[v];
            }
            if (a.lambda) a.lambda[mb.p] = lam;
            if (lam != 0.0) {  // lambda == 0 can never infect
              // the Philox counter word: in a mirrored supergroup a coalesced read of the slot-ordered ids
              const uint32_t gid = a.inject ? 0u : a.marr ? a.mgid[c.slot0 + midx] : jb_gid(a, mb.p);
              const int v = jb_draw_core<VM>(a, gid, lam_v, lam, ordinal);
              if (v >= 0) {
                uint32_t infector = 0xFFFFFFFFu;
                if (EV)
                  infector = jb_blame_group<VM>(a, c, mb.p, (int)mb.lsg(), v, Lg, years, n_years, n_l,
                                                T_w + (size_t)v * L, row_l[v * L + mb.lsg()], beta, dt);
                jb_emit(a, mb.p, (uint32_t)v, g, infector);
              }
            }
          }
        }
      }
    }
  }
  if (tg == 0) atomicAdd(&a.counters[CNT_DRAWS], nsus);
}

// EV: record infection events (who infected whom, JB_STEP_RECORD_EVENTS).  A separate
// instantiation, so that the rare blame path costs the common launch neither registers nor stack.
template <int VM, int WPG, bool EV>
__global__ void __launch_bounds__(JB_CTA) k_groups(const __grid_constant__ GroupArgs a) {
  jb_stamp(3); jb_stamp(4);
  constexpr int GPC = JB_CTA_WARPS / WPG;      // groups per CTA
  const uint32_t gs = (threadIdx.x >> 5) / WPG;
  if (!a.n_list_dev) {
    const uint32_t gi_ = blockIdx.x * GPC + gs;
    if (gi_ < a.n_list) jb_group_one<VM, WPG, EV>(a, gi_);  // uniform over the group's threads
    return;
  }
  // the list's length lives on the device (households being visited in this step): a fixed grid walks it
  const uint32_t n = min(*a.n_list_dev, a.n_list);
  for (uint32_t gi_ = blockIdx.x * GPC + gs; gi_ < n; gi_ += gridDim.x * GPC) {
    jb_group_one<VM, WPG, EV>(a, gi_);
    jb_grp_sync<WPG>();  // the group's shared memory is reused
  }
}

// ---- tile kernel ----------------------------------------------------------------------------
// Small groups (static members, contact matrix dim <= 4) are packed at world-create time into
// 32-slot tiles that never split a group.  STREAM: the tile's slots ARE the (renumbered) person
// indices, so the person bytes are read as a coalesced stream with no index indirection
// (households).  Otherwise the slot holds the person index and the byte is gathered (companies).
//
// Phase A (detection): one thread owns 16 consecutive slots (a lane pair owns a tile): 128-bit
// loads of the slot descriptors and person bytes, which are also parked in shared memory; per-tile
// bit masks are exchanged with one shuffle; the (few) groups that hold both a present infector and
// a present susceptible are appended to the block's work list in shared memory.  Everything else
// costs a handful of bit operations.
// Phase B (processing): one LANE per listed group.  The lane walks its group's slots in shared
// memory in slot order -- the reference's own left-to-right order (Python `sum`,
// interaction.py:463-467) -- gathers the infectors' transmission probabilities, and computes
// Lambda for its susceptibles.  Philox + exp are the expensive part: draws are queued per warp
// and executed 32 at a time on fully populated warps.
#define JB_TILE_THREADS 128
#define JB_TILE_WARPS (JB_TILE_THREADS / 32)
#define JB_TILE_SLOTS (JB_TILE_THREADS * 16)
#define JB_TILE_ITEMS (JB_TILE_SLOTS / 2)  // a listed group has >= 2 slots
#define JB_DRAW_QUEUE 64
// JB_TILE_PER_WARP = 1: every warp detects AND processes its own 512 slots (own work list, no block-wide barrier
// between the phases).  Measured (profiles/README.md, round 2): 51 % more warp instructions (every warp runs its own,
// two-thirds full, processing batch) for the same time at 56 M people and 6 % less at 7 M -- kept as an option.
#ifndef JB_TILE_PER_WARP
#define JB_TILE_PER_WARP 0
#endif


// TM: where a slot's person byte comes from.  JB_TM_GATHER: the slot holds a person index, the byte is gathered from
// `phot`; JB_TM_STREAM: slot == person index (households), `phot` is streamed; JB_TM_MIRROR: the supergroup owns a
// slot-ordered copy of its members' state bytes (`marr`, kept current by the placement), streamed like `phot`, and
// slot-ordered copies of the members' ids -- nothing is gathered per slot.
#define JB_TM_GATHER 0
#define JB_TM_STREAM 1
#define JB_TM_MIRROR 2
template <int TM, int VM, bool EV>
// (stream tiles are occupancy hungry: 48 registers measured 20 % faster than 70; the gather
// flavour keeps its registers for the 16 gathers in flight per thread)
#ifndef JB_TILE_MINB
#define JB_TILE_MINB 10
#endif
__global__ void __launch_bounds__(JB_TILE_THREADS, TM != JB_TM_GATHER ? JB_TILE_MINB : 1) k_tiles(const __grid_constant__ GroupArgs a) {
  constexpr bool STREAM = TM == JB_TM_STREAM;
  if (STREAM) jb_stamp(2);
  const double* const trans = jb_trans_read(a.trans, a.NP, a.sp);
  constexpr bool PW = JB_TILE_PER_WARP != 0;
  constexpr int ITEMS_W = JB_TILE_ITEMS / JB_TILE_WARPS;  // work-list entries of one warp (per-warp lists)
  __shared__ uint2 s_items[JB_TILE_ITEMS];  // {(block-local first slot << 5) | (len - 1), group}
  __shared__ uint4 s_hot4[JB_TILE_THREADS], s_info4[JB_TILE_THREADS];
  __shared__ uint32_t s_mem[TM == JB_TM_GATHER ? JB_TILE_SLOTS : 1];
  __shared__ double s_cm[16];
  __shared__ uint4 s_tm[JB_TILE_SLOTS / 32];     // per tile of the block: present | infector | subgroup bit 0 | subgroup bit 1 (slot masks)
  __shared__ uint32_t s_tgi[JB_TILE_SLOTS / 32];  // the group word shared by the groups of each of the block's tiles, or MIXED
  __shared__ double s_rcp[33];  // 1 / n for the subgroup sizes a tile can hold: the row sums multiply instead of dividing
  __shared__ uint32_t s_nw[JB_TILE_WARPS];  // work-list lengths (per-warp lists; [0] = the block's list otherwise)
  __shared__ double q_lam[JB_TILE_WARPS][VM][JB_DRAW_QUEUE];
  __shared__ uint32_t q_x[JB_TILE_WARPS][JB_DRAW_QUEUE], q_ord[JB_TILE_WARPS][JB_DRAW_QUEUE];  // block-local slot, ordinal
  __shared__ uint16_t q_item[JB_TILE_WARPS][JB_DRAW_QUEUE];  // work item of the queued draw (blame)
  const uint8_t* s_hot = reinterpret_cast<const uint8_t*>(s_hot4);
  const uint8_t* s_info = reinterpret_cast<const uint8_t*>(s_info4);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t& s_n = s_nw[PW ? wib : 0];
  uint2* const my_items = s_items + (PW ? wib * ITEMS_W : 0);
  if (PW ? lane == 0 : threadIdx.x == 0) s_n = 0;
  // contact matrix: requested now, parked in shared memory behind phase A (not a DRAM round trip
  // in front of the block's first barrier); per-warp lists: every warp parks the same 16 values
  const double cm_mine = (PW ? lane : (int)threadIdx.x) < a.dim * a.dim ? a.cm[PW ? lane : threadIdx.x] : 0.0;
  if (PW) __syncwarp(); else __syncthreads();

  // ---- phase A ---------------------------------------------------------------------------
  const uint32_t slot_base = blockIdx.x * JB_TILE_SLOTS;  // first slot of this block
  // person index of block-local slot x
  auto person = [&](uint32_t x) -> uint32_t {
    if (TM == JB_TM_STREAM) return a.stream_base + slot_base + x;
    if (TM == JB_TM_MIRROR) return a.tp_member[slot_base + x];
    return s_mem[x];
  };
  // global id (Philox counter word) of the person in block-local slot x
  auto gid_of = [&](uint32_t x) -> uint32_t {
    if (TM == JB_TM_MIRROR) return a.mgid[a.stream_base + slot_base + x];
    return jb_gid(a, person(x));
  };
  {
    const uint32_t n_chunks = a.n_tiles * 2;
    const uint32_t c = blockIdx.x * JB_TILE_THREADS + threadIdx.x;
    const bool ok = c < n_chunks;
    const uint32_t slot0 = c * 16;
    uint32_t valid = 0, first = 0, pres = 0, inf = 0, lg0 = 0, lg1 = 0;
    uint4 iv = make_uint4(0u, 0u, 0u, 0u), sv = make_uint4(0u, 0u, 0u, 0u);
    uint32_t tg0 = 0;
    // LEAN: what the slot descriptors say about a tile never changes -- which slots are valid, where the groups start,
    // the subgroup planes -- and is read as four words per tile (`tile_static`, built with the layout) instead of being
    // recomputed from 32 descriptor bytes by every launch; with one feeding activity for the whole supergroup (the
    // households' residence) only the persons' bytes are streamed and turned into the present / infector masks.
    const bool LEAN = STREAM && !EV && a.tile_static != nullptr;
    uint4 ts = make_uint4(0u, 0u, 0u, 0u);
    if (LEAN) {
      if (ok) {
        if (!(lane & 1)) {
          tg0 = a.tile_group0[c >> 1];
          s_tgi[threadIdx.x >> 1] = a.tile_ginfo ? a.tile_ginfo[c >> 1] : JB_TILE_GINFO_MIXED;
          ts = a.tile_static[c >> 1];
        }
        sv = *reinterpret_cast<const uint4*>(a.phot + a.stream_base + slot0);
        const uint32_t act4 = (uint32_t)a.uniform_act * 0x01010101u;
        const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t st4 = sw[q];
          const uint32_t pres4 = __vcmpeq4(st4 & 0x07070707u, act4) & 0x01010101u;
          const uint32_t inf4 = (st4 >> 3) & pres4;
          pres |= (((pres4 * 0x00204081u) >> 21) & 0xFu) << (4 * q);
          inf |= (((inf4 * 0x00204081u) >> 21) & 0xFu) << (4 * q);
        }
      }
    } else
    if (ok) {
      if (!(lane & 1)) {  // with the other loads, not behind the bit logic
        tg0 = a.tile_group0[c >> 1];
        s_tgi[threadIdx.x >> 1] = a.tile_ginfo ? a.tile_ginfo[c >> 1] : JB_TILE_GINFO_MIXED;
      }
      iv = *reinterpret_cast<const uint4*>(a.tp_info + slot0);
      if (TM == JB_TM_STREAM) {
        sv = *reinterpret_cast<const uint4*>(a.phot + a.stream_base + slot0);
      } else if (TM == JB_TM_MIRROR) {
        // state bytes -> hot bytes, four at a time: absent (any bit of the step's absence mask) -> no activity
        const uint4 mv = *reinterpret_cast<const uint4*>(a.marr + a.stream_base + slot0);
        const uint32_t abs4 = jb_mirror_absmask(a.sp->activity_mask, a.sp->flags) * 0x01010101u;
        const uint32_t act4 = (uint32_t)JB_ACT_PRIMARY * 0x01010101u  /* mirrored slots are fed by the primary activity */;
        auto cvt = [&](uint32_t w) -> uint32_t {
          const uint32_t ab = w & abs4;
          const uint32_t nz = ((((ab & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | ab) >> 7) & 0x01010101u;  // 1 where absent
          return (act4 ^ (nz * ((uint32_t)JB_ACT_PRIMARY ^ (uint32_t)JB_ACT_NONE))) | (w & 0x88888888u) |
                 ((w & 0x07070707u) << 4);
        };
        sv = make_uint4(cvt(mv.x), cvt(mv.y), cvt(mv.z), cvt(mv.w));
      } else {
        uint32_t mem[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 mv = *reinterpret_cast<const uint4*>(a.tp_member + slot0 + 4 * q);
          mem[4 * q] = mv.x; mem[4 * q + 1] = mv.y; mem[4 * q + 2] = mv.z; mem[4 * q + 3] = mv.w;
          *reinterpret_cast<uint4*>(&s_mem[threadIdx.x * 16 + 4 * q]) = mv;
        }
        uint32_t stg[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) stg[i] = mem[i] != 0xFFFFFFFFu ? a.phot[mem[i]] : 0u;
        sv.x = stg[0] | (stg[1] << 8) | (stg[2] << 16) | (stg[3] << 24);
        sv.y = stg[4] | (stg[5] << 8) | (stg[6] << 16) | (stg[7] << 24);
        sv.z = stg[8] | (stg[9] << 8) | (stg[10] << 16) | (stg[11] << 24);
        sv.w = stg[12] | (stg[13] << 8) | (stg[14] << 16) | (stg[15] << 24);
      }
      // four slots at a time: one byte-wise compare of the descriptors' activity against the
      // persons' activity
      const uint32_t iw[4] = {iv.x, iv.y, iv.z, iv.w}, sw[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t hi4 = iw[q], st4 = sw[q];
        const uint32_t valid4 = (hi4 >> 3) & 0x01010101u;
        const uint32_t first4 = (hi4 >> 4) & valid4;
        const uint32_t pres4 = __vcmpeq4(hi4 & 0x07070707u, st4 & 0x07070707u) & valid4;
        const uint32_t inf4 = (st4 >> 3) & pres4;
        // bit 0 of every byte -> 4-bit mask
        valid |= (((valid4 * 0x00204081u) >> 21) & 0xFu) << (4 * q);
        first |= (((first4 * 0x00204081u) >> 21) & 0xFu) << (4 * q);
        pres |= (((pres4 * 0x00204081u) >> 21) & 0xFu) << (4 * q);
        inf |= (((inf4 * 0x00204081u) >> 21) & 0xFu) << (4 * q);
        lg0 |= (((((hi4 >> 5) & 0x01010101u) * 0x00204081u) >> 21) & 0xFu) << (4 * q);  // JB_TI_LSG, bit planes
        lg1 |= (((((hi4 >> 6) & 0x01010101u) * 0x00204081u) >> 21) & 0xFu) << (4 * q);
      }
    }
    s_hot4[threadIdx.x] = sv;
    s_info4[threadIdx.x] = iv;
    // a tile = two neighbouring chunks: exchange the 16-bit masks inside the lane pair
    const uint32_t mine = first | (pres << 16);
    const uint32_t other = __shfl_xor_sync(JB_FULL, mine, 1);
    const uint32_t iv_o = __shfl_xor_sync(JB_FULL, inf | (valid << 16), 1);
    const uint32_t lg_o = LEAN ? 0u : __shfl_xor_sync(JB_FULL, lg0 | (lg1 << 16), 1);
    uint32_t draws = 0;
    const uint32_t t = c >> 1;
    if (ok && !(lane & 1)) {  // the even lane of the pair owns the tile
      // (LEAN: ts = {valid slots, first slots of the groups, subgroup plane 0, subgroup plane 1})
      const uint32_t first32 = LEAN ? ts.y : (first & 0xFFFFu) | (other << 16);
      const uint32_t pres32 = LEAN ? (pres | (other & 0xFFFF0000u)) & ts.x : pres | (other & 0xFFFF0000u);
      const uint32_t inf32 = LEAN ? (inf | (iv_o << 16)) & ts.x : inf | (iv_o << 16);
      const uint32_t n_valid = LEAN ? (uint32_t)__popc(ts.x) : 32 - __clz((valid & 0xFFFFu) | (iv_o & 0xFFFF0000u));  // valid slots are a prefix
      const uint32_t sus32 = pres32 & ~inf32;
      // the processing phase works on these masks instead of walking the slots (bit b = slot b of the tile)
      s_tm[threadIdx.x >> 1] = LEAN ? make_uint4(pres32, inf32, ts.z, ts.w)
                                    : make_uint4(pres32, inf32, lg0 | (lg_o << 16), lg1 | (lg_o & 0xFFFF0000u));
      uint32_t rem = inf32;
      while (rem) {
        const uint32_t src = __ffs(rem) - 1;
        const uint32_t le = 0xFFFFFFFFu >> (31 - src);
        const uint32_t below = first32 & le, above = first32 & ~le;
        const uint32_t start = 31 - __clz(below | 1u);
        const uint32_t end = above ? (uint32_t)__ffs(above) - 1u : n_valid;
        const uint32_t gmask = (end == 32u ? 0xFFFFFFFFu : ((1u << end) - 1u)) & ~((1u << start) - 1u);
        rem &= ~gmask;
        const uint32_t nsus = __popc(gmask & sus32);
        if (!nsus) continue;  // must_timestep needs a susceptible too (interactive.py:136)
        const uint32_t g = tg0 + __popc(below) - 1u;
        if (a.vslot && a.vslot[g - a.g0] != JB_VIS_NONE) continue;  // being visited: the visited-list launch has it
        if (a.mode == 1) { a.draw_cnt[g] = nsus; continue; }
        draws += nsus;
        const uint32_t pos = atomicAdd(&s_n, 1u);
        my_items[pos] = make_uint2(((t * 32 + start - slot_base) << 5) | (end - start - 1), g);
      }
    }
    draws = __reduce_add_sync(JB_FULL, draws);
    if (lane == 0 && draws) atomicAdd(&a.counters[CNT_DRAWS], draws);
  }
  if ((PW ? lane : (int)threadIdx.x) < a.dim * a.dim) s_cm[PW ? lane : threadIdx.x] = cm_mine;
  {
    const int k = PW ? lane : (int)threadIdx.x;  // (per-warp lists: every warp parks the same values)
    if (k < 32) s_rcp[k + 1] = 1.0 / (double)(k + 1);
  }
  if (PW) __syncwarp(); else __syncthreads();
  const uint32_t n_work = s_n;
  if (STREAM) jb_stamp_max(0);
  if (n_work == 0 || a.mode == 2) { if (STREAM) jb_stamp_max(1); return; }  // mode 2: detection only (profiling aid)

  // ---- phase B ---------------------------------------------------------------------------
  const int V = (VM > 1) ? a.V : 1;
  const int dim = a.dim;
  const uint32_t lane_lt = (1u << lane) - 1u;
  const double dt = a.sp->dt;
  uint32_t qn = 0;  // warp-uniform queue length

  auto drain = [&](uint32_t n) {
    if ((uint32_t)lane < n) {
      double lam_v[VM];
      double lam = 0.0;
#pragma unroll
      for (int v = 0; v < VM; ++v) { lam_v[v] = q_lam[wib][v][lane]; lam += lam_v[v]; }
      const uint32_t x = q_x[wib][lane];
      const int v = jb_draw_core<VM>(a, a.inject ? 0u : gid_of(x), lam_v, lam, q_ord[wib][lane]);
      if (v >= 0) {
        const uint32_t p = person(x);
        const uint2 it = my_items[q_item[wib][lane]];
        uint32_t infector = 0xFFFFFFFFu;
        if (EV) {
          // rare path: rebuild the group's sums from the block's shared-memory copy of the slots,
          // then _blame_subgroup / _blame_individuals (interaction.py:643-662)
          const uint32_t loc = it.x >> 5, len = (it.x & 31u) + 1u;
          uint32_t npk = 0, my_i = 0;
          double Tj[4] = {0.0, 0.0, 0.0, 0.0};
          for (uint32_t m = 0; m < len; ++m) {
            const uint32_t info = s_info[loc + m], hot = s_hot[loc + m];
            if (((hot ^ info) & JB_PS_ACT_MASK) != 0u) continue;
            const uint32_t q = person(loc + m);
            const int l = (int)JB_TI_LSG(info);
            npk += 1u << (8 * l);
            if (q == p) my_i = (uint32_t)l;
            if ((hot & JB_PS_INFECTED) && (VM == 1 || (int)JB_HOT_VAR(hot) == v)) {
              const double t = trans[q];
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j == l) Tj[j] += t;
            }
          }
          const uint32_t gi = a.grp_info[it.y];
          const double beta = a.beta[JB_GINFO_BETA(gi) * a.R + JB_GINFO_REGION(gi)] * a.beta_scale[JB_GINFO_SCALE(gi)];
          double term[4], total = 0.0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            term[j] = 0.0;
            if (j < dim && Tj[j] != 0.0) {
              const int nj = (int)((npk >> (8 * j)) & 0xFFu);
              const int nn = ((int)my_i == j) ? max(1, nj - 1) : nj;
              term[j] = s_cm[my_i * dim + j] * Tj[j] / (double)nn * beta * dt;
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
          for (uint32_t m = 0; m < len; ++m) {
            const uint32_t info = s_info[loc + m], hot = s_hot[loc + m];
            if (((hot ^ info) & JB_PS_ACT_MASK) != 0u || !(hot & JB_PS_INFECTED) || (int)JB_TI_LSG(info) != jb) continue;
            if (VM > 1 && (int)JB_HOT_VAR(hot) != v) continue;
            const uint32_t q = person(loc + m);
            c2 += trans[q];
            infector = q;
            if (c2 > target2) break;
          }
        }
        jb_emit(a, p, (uint32_t)v, it.y, infector);
      }
    }
    __syncwarp();
  };

  for (uint32_t wbase = PW ? 0u : wib * 32u; wbase < n_work; wbase += PW ? 32u : JB_TILE_WARPS * 32u) {
    const uint32_t w = wbase + lane;
    const bool have_item = w < n_work;
    const uint2 item = have_item ? my_items[w] : make_uint2(0u, 0u);
    const uint32_t loc = item.x >> 5, len = have_item ? (item.x & 31u) + 1u : 0u, g = item.y;
    // The group as slot masks of its tile (bit b = slot b): present, infectors, susceptibles, and the members of
    // each subgroup.  Sizes are popcounts; only the infectors and the susceptibles are walked (set bits, in slot
    // order -- the reference's left-to-right order) instead of every slot three times.
    const uint32_t tile0 = loc & ~31u;  // block-local first slot of the group's tile
    const uint4 tm = have_item ? s_tm[loc >> 5] : make_uint4(0u, 0u, 0u, 0u);
    const uint32_t gm = have_item ? ((len >= 32u ? 0xFFFFFFFFu : ((1u << len) - 1u)) << (loc & 31u)) : 0u;
    const uint32_t P = tm.x & gm, I = tm.y & gm;
    uint32_t S = P & ~I;
    // Everything phase B gathers per group (the infectors' transmission probabilities, the group
    // word, the members' global ids for the Philox counter) is independent: prefetch it all into
    // L2 first so that the DRAM latencies overlap instead of chaining.
    const uint32_t tgi = have_item ? s_tgi[loc >> 5] : 0u;
    if (have_item) {
      if (tgi == JB_TILE_GINFO_MIXED) jb_prefetch_l2(a.grp_info + g);
      if (TM == JB_TM_MIRROR) {
        jb_prefetch_l2(a.mgid + a.stream_base + slot_base + loc);
      } else {
        const uint32_t pf = person(loc);
        if (pf < a.N) jb_prefetch_l2(a.int2ext + pf);
      }
    }
    for (uint32_t r = I; r; r &= r - 1u) jb_prefetch_l2(trans + person(tile0 + (uint32_t)__ffs(r) - 1u));
    double beta = 0.0;
    uint32_t ord = 0;
    if (have_item) {
      const uint32_t gi = tgi != JB_TILE_GINFO_MIXED ? tgi : a.grp_info[g];
      beta = a.beta[JB_GINFO_BETA(gi) * a.R + JB_GINFO_REGION(gi)] * a.beta_scale[JB_GINFO_SCALE(gi)];
      ord = a.inject ? a.draw_base[g] : 0u;
    }
    // subgroup sizes (four 8-bit counters) and summed transmission probabilities
    const uint32_t npk = (uint32_t)__popc(P & ~tm.w & ~tm.z) | ((uint32_t)__popc(P & ~tm.w & tm.z) << 8) |
                         ((uint32_t)__popc(P & tm.w & ~tm.z) << 16) | ((uint32_t)__popc(P & tm.w & tm.z) << 24);
    double T[VM][4];
#pragma unroll
    for (int v = 0; v < VM; ++v)
#pragma unroll
      for (int k = 0; k < 4; ++k) T[v][k] = 0.0;
    for (uint32_t r = I; r; r &= r - 1u) {
      const uint32_t bit = (uint32_t)__ffs(r) - 1u;
      const int l = (int)(((tm.z >> bit) & 1u) | (((tm.w >> bit) & 1u) << 1));
      const double t = trans[person(tile0 + bit)];
      const int pv = (VM > 1) ? (int)JB_HOT_VAR((uint32_t)s_hot[tile0 + bit]) : 0;
#pragma unroll
      for (int v = 0; v < VM; ++v)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (v == pv && j == l) T[v][j] += t;
    }
    // Lambda of every susceptible (interaction.py:594-631), draws queued
    while (__any_sync(JB_FULL, S != 0u)) {
      const bool sus = S != 0u;
      const uint32_t bit = sus ? (uint32_t)__ffs(S) - 1u : 0u;
      S &= S - 1u;
      const uint32_t x = tile0 + bit;  // block-local slot
      double lam_v[VM];
      double lam = 0.0;
#pragma unroll
      for (int v = 0; v < VM; ++v) lam_v[v] = 0.0;
      if (sus) {
        const uint32_t hot = s_hot[x];
        // the person index is only needed for a susceptibility that is neither 0 nor 1 (and in test modes)
        const uint32_t p = (TM != JB_TM_MIRROR || VM > 1 || a.lambda || JB_HOT_CODE(hot) == JB_SC_READ) ? person(x) : 0u;
        const int i = (int)(((tm.z >> bit) & 1u) | (((tm.w >> bit) & 1u) << 1));
#pragma unroll
        for (int v = 0; v < VM; ++v) {
          double r = 0.0;
          if (v < V) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < dim && T[v][j] != 0.0) {
                const int nj = (int)((npk >> (8 * j)) & 0xFFu);
                const int nn = (i == j) ? max(1, nj - 1) : nj;  // interaction.py:469-471
                // (times 1/nn from the table: the fp64 division was a seventh of the kernel's instructions; the
                // quotient may differ from the reference's in the last bit)
                r += s_cm[i * dim + j] * T[v][j] * s_rcp[nn] * beta * dt;
              }
            r *= jb_susc<VM>(a, p, hot, v);
          }
          lam_v[v] = r;
          lam += r;
        }
        if (a.lambda) a.lambda[p] = lam;
      }
      // queue the draws that can succeed (lambda == 0 never infects)
      const bool want = sus && lam != 0.0;
      const uint32_t wm = __ballot_sync(JB_FULL, want);
      if (wm) {
        if (want) {
          const uint32_t pos = qn + __popc(wm & lane_lt);
          q_x[wib][pos] = x;
          q_ord[wib][pos] = ord;
          if (EV) q_item[wib][pos] = (uint16_t)w;
#pragma unroll
          for (int v = 0; v < VM; ++v) q_lam[wib][v][pos] = lam_v[v];
        }
        qn += __popc(wm);
        __syncwarp();
        if (qn >= 32) {
          drain(32);
          const uint32_t rest = qn - 32;
          uint32_t mp = 0, mo = 0, mi = 0;
          double ml[VM];
          if ((uint32_t)lane < rest) {
            mp = q_x[wib][32 + lane]; mo = q_ord[wib][32 + lane]; mi = q_item[wib][32 + lane];
#pragma unroll
            for (int v = 0; v < VM; ++v) ml[v] = q_lam[wib][v][32 + lane];
          }
          __syncwarp();
          if ((uint32_t)lane < rest) {
            q_x[wib][lane] = mp; q_ord[wib][lane] = mo; q_item[wib][lane] = (uint16_t)mi;
#pragma unroll
            for (int v = 0; v < VM; ++v) q_lam[wib][v][lane] = ml[v];
          }
          __syncwarp();
          qn = rest;
        }
      }
      if (sus) ++ord;
    }
  }
  if (qn) drain(qn);
  if (STREAM) jb_stamp_max(1);
}
