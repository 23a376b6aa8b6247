// Hand-written sm_100a kernels of the JUNE per-timestep hot path.
//
//   k_place          ActivityManager.move_people_to_active_subgroups   (activity_manager.py:255-402)
//   k_groups_thread  Interaction.time_step_for_group, small groups     (interaction.py:516-641)
//   k_groups_warp    same, one warp per group (companies, schools, hospitals, care homes, ...)
//   k_newinf         InfectionSelector._make_infection                 (infection_selector.py:146-332)
//   k_health         Epidemiology.update_health_status                 (epidemiology.py:217-303)
//   k_halo_pack/unpack, k_halo_return   MovablePeople exchange         (mpi_wrapper.py:145-293)
//
// All of them are HBM-bound streams / gathers of byte, int and fp64 data: there is no dense
// contraction anywhere (largest "matrix" is a <=60x60 lookup), so no tensor cores are used.
#pragma once
#include <math.h>

#include "jb_internal.cuh"
#include "jb_rng.cuh"

#define JB_FULL 0xFFFFFFFFu

// ------------------------------------------------------------------------------------------
// Stage A: placement.  The reference walks the activity hierarchy per person and takes the first
// activity with a non-None subgroup (activity_manager.py:380-398); stay-home policies reduce the
// activity list to (medical_facility, residence) (individual_policies.py:87-119).  Here the
// chosen activity is written into the low bits of the person's state byte; fixed-membership
// subgroups test that byte instead of being scattered into.  People whose subgroup is dynamic
// (hospital patients) are scattered into the per-group bin of their ward.
// 16 persons per thread (one 128-bit load of each byte array), grid-stride, counters reduced in
// registers -> shared memory -> one global atomic per block and counter.
// ------------------------------------------------------------------------------------------
struct DynRanges {
  uint32_t n;
  uint32_t first[4], count[4], bin_base[4];
};

struct PlaceArgs {
  uint8_t* pstate;
  const uint8_t* phas;
  const int32_t* pmed;
  uint32_t* dyn_cnt;   // per dynamic group: members appended this step
  uint32_t* dyn_bin;   // [n_bins][bin_cap] entries (lsg << 28) | person
  uint32_t bin_cap;
  DynRanges dr;
  uint32_t* counters;
  uint32_t N;
  const StepDev* sp;
};

// activity chosen for a person byte: pure function of (dead, severe, has-medical, fixed
// activities) and of the step's activity mask / policy flags -> 512-entry table per block
__device__ __forceinline__ uint32_t jb_place_key(uint32_t st, uint32_t has) {
  return (((st >> 3) & 0xFu) << 5) | (has & 0x1Fu);  // bits: INF DEAD SEVERE MED | has[5]
}

__global__ void __launch_bounds__(256) k_place(const __grid_constant__ PlaceArgs a) {
  __shared__ uint8_t s_act[512];      // 0..4 activity, 7 none (dead), 15 alive without any activity
  __shared__ uint32_t s_cnt[8];
  const uint32_t mask = a.sp->activity_mask, flags = a.sp->flags;
  for (uint32_t k = threadIdx.x; k < 512; k += blockDim.x) {
    const uint32_t bits = k >> 5, has0 = k & 0x1Fu;
    const bool dead = bits & 2u, severe = bits & 4u, med = bits & 8u;
    uint32_t act = JB_ACT_NONE;
    if (!dead) {
      uint32_t allowed = mask;
      if ((flags & JB_STEP_SEVERE_STAY_HOME) && severe) allowed &= (1u << JB_ACT_MEDICAL) | (1u << JB_ACT_RESIDENCE);
      const uint32_t cand = allowed & (has0 | (med ? 1u << JB_ACT_MEDICAL : 0u));
      act = cand ? (uint32_t)__ffs(cand) - 1u : 15u;  // hierarchy order == bit order
    }
    s_act[k] = (uint8_t)act;
  }
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  // eight 8-bit counters (one per activity code) packed in a 64-bit register
  unsigned long long cnt = 0ull;
  uint32_t noact = 0;
  const uint32_t n_chunks = (a.N + 15) / 16;
  auto one = [&](uint32_t p, uint32_t st, uint32_t has) -> uint32_t {
    uint32_t act = s_act[jb_place_key(st, has)];
    if (act == 15u) { ++noact; act = JB_ACT_NONE; }
    cnt += 1ull << (8 * act);
    if (act == JB_ACT_MEDICAL) {  // rare: scatter the patient into its ward's bin
      const uint32_t med = (uint32_t)a.pmed[p];  // (hospital group << 8) | local subgroup
      const uint32_t g = med >> 8;
      bool okb = false;
      for (uint32_t r = 0; r < a.dr.n; ++r) {
        if (g - a.dr.first[r] < a.dr.count[r]) {
          const uint32_t bin = a.dr.bin_base[r] + (g - a.dr.first[r]);
          const uint32_t i = atomicAdd(&a.dyn_cnt[bin], 1u);
          if (i < a.bin_cap) { a.dyn_bin[(size_t)bin * a.bin_cap + i] = ((med & 0xFu) << 28) | p; okb = true; }
          break;
        }
      }
      if (!okb) atomicAdd(&a.counters[CNT_OVERFLOW], 1u);
    }
    return (st & ~JB_PS_ACT_MASK) | act;
  };
  uint32_t iters = 0;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += gridDim.x * blockDim.x) {
    const uint32_t p0 = c * 16;
    if (p0 + 16 <= a.N) {
      uint4 sv = *reinterpret_cast<const uint4*>(a.pstate + p0);
      const uint4 hv = *reinterpret_cast<const uint4*>(a.phas + p0);
      uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
      const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t o = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b)
          o |= one(p0 + q * 4 + b, (sw[q] >> (8 * b)) & 0xFFu, (hw[q] >> (8 * b)) & 0xFFu) << (8 * b);
        sw[q] = o;
      }
      *reinterpret_cast<uint4*>(a.pstate + p0) = make_uint4(sw[0], sw[1], sw[2], sw[3]);
    } else {
      for (uint32_t p = p0; p < a.N; ++p) a.pstate[p] = (uint8_t)one(p, a.pstate[p], a.phas[p]);
    }
    if (++iters == 15) {  // 15 * 16 = 240 < 256: flush the packed 8-bit counters
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&s_cnt[k], (uint32_t)((cnt >> (8 * k)) & 0xFFull));
      cnt = 0ull; iters = 0;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t v = __reduce_add_sync(JB_FULL, (uint32_t)((cnt >> (8 * k)) & 0xFFull));
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[k], v);
  }
  noact = __reduce_add_sync(JB_FULL, noact);
  if ((threadIdx.x & 31) == 0 && noact) atomicAdd(&a.counters[CNT_NOACT], noact);
  __syncthreads();
  if (threadIdx.x < JB_N_ACT && s_cnt[threadIdx.x]) atomicAdd(&a.counters[CNT_ACT0 + threadIdx.x], s_cnt[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------
// Stage B: interaction.
// ------------------------------------------------------------------------------------------
// Global person id (the Philox counter word): residents are renumbered internally
// (int2ext maps back to the caller's index), halo persons carry their home domain's id.
__device__ __forceinline__ uint32_t jb_gid(const GroupArgs& a, uint32_t p) {
  return p < a.N ? a.gid_base + a.int2ext[p] : a.halo_gid[p - a.N];
}

// Append one new infection.
__device__ __forceinline__ void jb_emit(const GroupArgs& a, uint32_t p, uint32_t v) {
  uint32_t i = atomicAdd(&a.counters[CNT_NEWINF], 1u);
  if (i < a.newinf_cap) {
    a.newinf_member[i] = p;
    a.newinf_var[i] = (uint8_t)v;
  } else {
    atomicAdd(&a.counters[CNT_OVERFLOW], 1u);
  }
}

// Bernoulli draw + variant choice for one susceptible (interaction.py:610-641).
template <int VM>
__device__ __forceinline__ void jb_draw_core(const GroupArgs& a, uint32_t p, const double* lam_v, double lam,
                                             uint32_t ordinal) {
  double u;
  if (a.inject) {
    u = ordinal < a.sp->n_inject ? a.inject[ordinal] : 2.0;
  } else {
    u = jb_bernoulli_uniform(a.sp->seed, a.sp->step, jb_gid(a, p));
  }
  if (u < 1.0 - exp(-lam)) {
    uint32_t v = 0;
    if (VM > 1 && a.V > 1) {
      // np.random.choice(ids, p=lam_v/lam): inverse CDF on a second, independent uniform
      double u2 = jb_variant_uniform(a.sp->seed, a.sp->step, jb_gid(a, p)) * lam;
      double c = 0.0;
      v = a.V - 1;
      for (int k = 0; k < a.V; ++k) {
        c += lam_v[k];
        if (u2 < c) { v = k; break; }
      }
    }
    jb_emit(a, p, v);
  }
}

template <int VM>
__device__ __forceinline__ void jb_draw(const GroupArgs& a, uint32_t p, const double* lam_v, double lam,
                                        uint32_t ordinal) {
  if (a.lambda) a.lambda[p] = lam;
  jb_draw_core<VM>(a, p, lam_v, lam, ordinal);
}

// ---- warp-per-group kernel ---------------------------------------------------------------
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

struct JbMember {
  uint32_t p, lsg;
  bool present, infected;
};

// Member `idx` of group (static registered slots first, then the sorted dynamic members, which
// the warp has staged in shared memory as (lsg << 28) | person).
__device__ __forceinline__ JbMember jb_load_member(const GroupArgs& a, uint32_t idx, uint32_t b, uint32_t ns,
                                                   const uint32_t* dyn_s, uint32_t ntot) {
  JbMember mbr;
  mbr.p = 0; mbr.lsg = 0xFFFFu; mbr.present = false; mbr.infected = false;
  if (idx >= ntot) return mbr;
  uint32_t st;
  if (idx < ns) {
    uint32_t m = b + idx;
    mbr.p = a.reg_member[m];
    uint32_t info = a.reg_info[m];
    st = a.pstate[mbr.p];
    mbr.lsg = JB_INFO_LSG(info);
    mbr.present = (st & JB_PS_ACT_MASK) == JB_INFO_ACT(info);
  } else {
    uint32_t k = dyn_s[idx - ns];
    mbr.p = k & 0x0FFFFFFFu;
    mbr.lsg = k >> 28;
    st = a.pstate[mbr.p];
    mbr.present = true;
  }
  mbr.infected = (st & JB_PS_INFECTED) != 0;
  return mbr;
}

// Shared memory per warp: n[L] u32 | sus_static[L] u32 | sus_dyn[L] u32 | pad | T[V*L] f64 | row[V*L] f64
// | dynamic member staging (bin_cap u32, only for supergroups with dynamic members)
__host__ __device__ inline size_t jb_warp_smem_bytes(int L, int V, uint32_t bin_cap) {
  size_t ints = ((size_t)3 * L + 1) & ~(size_t)1;
  return ints * 4 + (size_t)2 * V * L * 8 + (size_t)bin_cap * 4;
}

template <int VM>
__global__ void __launch_bounds__(128) k_groups_warp(const __grid_constant__ GroupArgs a) {
  extern __shared__ __align__(8) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t gi_ = blockIdx.x * (blockDim.x >> 5) + wib;
  if (gi_ >= a.n_list) return;
  const uint32_t g = a.group_list ? a.group_list[gi_] : a.g0 + gi_;
  const int L = a.L, V = (VM > 1) ? a.V : 1;
  const uint32_t bin_cap = a.dyn_cnt ? a.bin_cap : 0u;
  unsigned char* base = smem_raw + (size_t)wib * jb_warp_smem_bytes(L, V, bin_cap);
  uint32_t* n_l = (uint32_t*)base;
  uint32_t* ss_l = n_l + L;   // susceptibles among static members, per subgroup
  uint32_t* sd_l = ss_l + L;  // susceptibles among dynamic members
  double* T_l = (double*)(base + ((((size_t)3 * L + 1) & ~(size_t)1) * 4));
  double* row_l = T_l + (size_t)V * L;
  uint32_t* dyn_s = (uint32_t*)(row_l + (size_t)V * L);

  const uint32_t b = a.grp_off[g], ns = a.grp_off[g + 1] - b;
  uint32_t nd = 0;
  if (a.dyn_cnt) {
    // stage this group's bin and sort it: members arrive in arbitrary (atomic) order, the
    // reference order is (subgroup, person) -- bitonic sort by the warp in shared memory
    const uint32_t bin = a.dyn_bin_base + (g - a.g0);
    nd = min(a.dyn_cnt[bin], bin_cap);
    if (nd) {
      uint32_t npad = 32;
      while (npad < nd) npad <<= 1;
      const uint32_t* src = a.dyn_bin + (size_t)bin * bin_cap;
      for (uint32_t i = lane; i < npad; i += 32) dyn_s[i] = i < nd ? src[i] : 0xFFFFFFFFu;
      __syncwarp();
      for (uint32_t k = 2; k <= npad; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          for (uint32_t i = lane; i < npad; i += 32) {
            const uint32_t ixj = i ^ j;
            if (ixj > i) {
              const uint32_t x = dyn_s[i], y = dyn_s[ixj];
              const bool asc = (i & k) == 0;
              if ((x > y) == asc) { dyn_s[i] = y; dyn_s[ixj] = x; }
            }
          }
          __syncwarp();
        }
    }
  }
  const uint32_t ntot = ns + nd;
  if (ntot == 0) {
    if (a.mode == 1 && lane == 0) a.draw_cnt[g] = 0;
    return;
  }
  // number of local subgroups of this group
  int Lg = a.dim;
  const uint8_t* years = nullptr;
  int n_years = 0;
  if (a.kind == JB_KIND_SCHOOL) {
    uint32_t aux = a.grp_aux[g];
    n_years = aux & 0xFFu;
    years = a.school_years + (aux >> 8);
    Lg = n_years + 1;
  }
  for (int i = lane; i < Lg; i += 32) {
    n_l[i] = 0; ss_l[i] = 0; sd_l[i] = 0;
    for (int v = 0; v < V; ++v) T_l[v * L + i] = 0.0;
  }
  __syncwarp();

  // ---- pass 1: subgroup sizes and summed transmission probabilities ----------------------
  // Static and dynamic members are walked in separate batches so that the subgroup keys stay
  // sorted inside every 32-wide batch (needed by the segmented scan).
  uint32_t nsus = 0, ninf = 0;
  constexpr int UB = 4;  // batches in flight: the member -> state -> value chain is latency bound
  for (int phase = 0; phase < 2; ++phase) {
    const uint32_t lo = phase ? ns : 0u, hi = phase ? ntot : ns;
    for (uint32_t it0 = lo; it0 < hi; it0 += 32 * UB) {
      JbMember mbs[UB];
      double tvs[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const uint32_t idx = it0 + 32 * u + lane;
        mbs[u] = jb_load_member(a, idx < hi ? idx : ntot, b, ns, dyn_s, ntot);
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) tvs[u] = (mbs[u].present && mbs[u].infected) ? a.trans[mbs[u].p] : 0.0;
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        if (it0 + 32 * u >= hi) break;
        const JbMember mb = mbs[u];
        const uint32_t key = mb.lsg;
        const uint32_t sus = (mb.present && !mb.infected) ? 1u : 0u;
        uint32_t packed = (mb.present ? 1u : 0u) | (sus << 16);
        packed = jb_seg_scan<uint32_t>(packed, key, lane);
        const uint32_t knext = __shfl_down_sync(JB_FULL, key, 1);
        const bool tail = (lane == 31 || knext != key) && key < (uint32_t)Lg;
        if (tail) {
          n_l[key] += packed & 0xFFFFu;
          if (phase == 0) ss_l[key] += packed >> 16; else sd_l[key] += packed >> 16;
        }
        nsus += __popc(__ballot_sync(JB_FULL, sus));
        const bool inf = mb.present && mb.infected;
        const uint32_t infm = __ballot_sync(JB_FULL, inf);
        ninf += __popc(infm);
        if (infm) {
          const int pv = (VM > 1 && inf) ? a.pvar[mb.p] : 0;
          for (int v = 0; v < V; ++v) {
            double tv = (pv == v) ? tvs[u] : 0.0;
            tv = jb_seg_scan<double>(tv, key, lane);
            if (tail) T_l[v * L + key] += tv;
          }
        }
        __syncwarp();
      }
    }
  }
  const bool must = nsus > 0 && ninf > 0;
  if (a.mode == 1) {
    if (lane == 0) a.draw_cnt[g] = must ? nsus : 0;
    return;
  }
  if (!must) return;

  // ---- infector tensor row sums (interaction.py:451-478, :619) ---------------------------
  const uint32_t gi = a.grp_info[g];
  const double beta = a.beta[JB_GINFO_BETA(gi) * a.R + JB_GINFO_REGION(gi)] * a.beta_scale[JB_GINFO_SCALE(gi)];
  for (int i = lane; i < Lg; i += 32) {
    for (int v = 0; v < V; ++v) {
      double s = 0.0;
      for (int j = 0; j < Lg; ++j) {
        double t = T_l[v * L + j];
        if (t != 0.0) {
          int nj = (int)n_l[j];
          int nn = (i == j) ? max(1, nj - 1) : nj;
          s += jb_cm_elem(a, i, j, years, n_years) * t / (double)nn * beta * a.sp->dt;
        }
      }
      row_l[v * L + i] = s;
    }
  }
  __syncwarp();
  // injected-uniform ordinals: susceptibles are numbered subgroup by subgroup, static members
  // before dynamic ones (interactive.py:86-131 dict insertion order)
  uint32_t ord_base = 0;
  if (a.inject) {
    ord_base = a.draw_base[g];
    if (lane == 0) {
      uint32_t run = 0;
      for (int i = 0; i < Lg; ++i) {
        uint32_t s_ = ss_l[i], d_ = sd_l[i];
        ss_l[i] = run; run += s_;
        sd_l[i] = run; run += d_;
      }
    }
    __syncwarp();
  }

  // ---- pass 2: one Bernoulli draw per susceptible ----------------------------------------
  for (int phase = 0; phase < 2; ++phase) {
    const uint32_t lo = phase ? ns : 0u, hi = phase ? ntot : ns;
    uint32_t* run = phase ? sd_l : ss_l;
    for (uint32_t it0 = lo; it0 < hi; it0 += 32 * UB) {
      JbMember mbs[UB];
      double s0[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const uint32_t idx = it0 + 32 * u + lane;
        mbs[u] = jb_load_member(a, idx < hi ? idx : ntot, b, ns, dyn_s, ntot);
      }
#pragma unroll
      for (int u = 0; u < UB; ++u)
        s0[u] = (mbs[u].present && !mbs[u].infected && mbs[u].lsg < (uint32_t)Lg) ? a.susc[mbs[u].p] : 0.0;
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        if (it0 + 32 * u >= hi) break;
        const JbMember mb = mbs[u];
        const bool sus = mb.present && !mb.infected && mb.lsg < (uint32_t)Lg;
        uint32_t ordinal = 0;
        if (a.inject) {
          const uint32_t key = mb.lsg;
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
            if (v < V) { r = row_l[v * L + mb.lsg]; sv = (v == 0) ? s0[u] : a.susc[(size_t)v * a.NP + mb.p]; }
            lam_v[v] = r * sv;
            lam += lam_v[v];
          }
          if (a.lambda) a.lambda[mb.p] = lam;
          if (lam != 0.0) jb_draw_core<VM>(a, mb.p, lam_v, lam, ordinal);  // lambda == 0 can never infect
        }
      }
    }
  }
  if (lane == 0) atomicAdd(&a.counters[CNT_DRAWS], nsus);
}

// ---- warp-per-group kernel, 1x1 contact matrix, static members (big companies, venues) ------
// One subgroup: the group needs only its size and the summed transmission probability, so the
// segmented scans / shared memory of the general kernel are replaced by ballots and one
// butterfly reduction.
template <int VM>
__global__ void __launch_bounds__(128) k_groups_warp_d1(const __grid_constant__ GroupArgs a) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t gi_ = blockIdx.x * (blockDim.x >> 5) + wib;
  if (gi_ >= a.n_list) return;
  const uint32_t g = a.group_list ? a.group_list[gi_] : a.g0 + gi_;
  const int V = (VM > 1) ? a.V : 1;
  const uint32_t b = a.grp_off[g], ns = a.grp_off[g + 1] - b;
  constexpr int UB = 4;
  const uint32_t lane_lt = (1u << lane) - 1u;
  uint32_t n = 0, nsus = 0, ninf = 0;
  double T[VM];
#pragma unroll
  for (int v = 0; v < VM; ++v) T[v] = 0.0;
  // the first 128 members stay in registers for the draw pass (most groups are that small)
  uint32_t p0[UB], st0[UB];
  bool pres0[UB];
  for (uint32_t it0 = 0; it0 < ns; it0 += 32 * UB) {
    uint32_t p[UB], st[UB];
    bool pres[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const uint32_t idx = it0 + 32 * u + lane;
      p[u] = idx < ns ? a.reg_member[b + idx] : 0xFFFFFFFFu;
      st[u] = idx < ns ? a.reg_info[b + idx] : 0u;  // info for now
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const uint32_t info = st[u];
      st[u] = p[u] != 0xFFFFFFFFu ? a.pstate[p[u]] : 0u;
      pres[u] = p[u] != 0xFFFFFFFFu && (st[u] & JB_PS_ACT_MASK) == JB_INFO_ACT(info);
    }
    if (it0 == 0) {
#pragma unroll
      for (int u = 0; u < UB; ++u) { p0[u] = p[u]; st0[u] = st[u]; pres0[u] = pres[u]; }
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const bool inf = pres[u] && (st[u] & JB_PS_INFECTED);
      n += __popc(__ballot_sync(JB_FULL, pres[u]));
      nsus += __popc(__ballot_sync(JB_FULL, pres[u] && !inf));
      ninf += __popc(__ballot_sync(JB_FULL, inf));
      if (inf) {
        const double t = a.trans[p[u]];
        const int pv = (VM > 1) ? a.pvar[p[u]] : 0;
#pragma unroll
        for (int v = 0; v < VM; ++v)
          if (v == pv) T[v] += t;
      }
    }
  }
  const bool must = nsus > 0 && ninf > 0;  // interactive.py:136
  if (a.mode == 1) {
    if (lane == 0) a.draw_cnt[g] = must ? nsus : 0;
    return;
  }
  if (!must) return;
#pragma unroll
  for (int v = 0; v < VM; ++v)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) T[v] += __shfl_xor_sync(JB_FULL, T[v], off);
  const uint32_t gi = a.grp_info[g];
  const double beta = a.beta[JB_GINFO_BETA(gi) * a.R + JB_GINFO_REGION(gi)] * a.beta_scale[JB_GINFO_SCALE(gi)];
  double row[VM];
#pragma unroll
  for (int v = 0; v < VM; ++v)
    row[v] = (v < V && T[v] != 0.0) ? a.cm[0] * T[v] / (double)max(1, (int)n - 1) * beta * a.sp->dt : 0.0;  // interaction.py:469-471
  uint32_t ord = a.inject ? a.draw_base[g] : 0u;
  for (uint32_t it0 = 0; it0 < ns; it0 += 32 * UB) {
    uint32_t p[UB];
    bool sus[UB];
    double s0[UB];
    if (it0 == 0) {
#pragma unroll
      for (int u = 0; u < UB; ++u) { p[u] = p0[u]; sus[u] = pres0[u] && !(st0[u] & JB_PS_INFECTED); }
    } else {
      uint32_t st[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const uint32_t idx = it0 + 32 * u + lane;
        p[u] = idx < ns ? a.reg_member[b + idx] : 0xFFFFFFFFu;
        st[u] = idx < ns ? a.reg_info[b + idx] : 0u;
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const uint32_t info = st[u];
        st[u] = p[u] != 0xFFFFFFFFu ? a.pstate[p[u]] : 0u;
        sus[u] = p[u] != 0xFFFFFFFFu && (st[u] & JB_PS_ACT_MASK) == JB_INFO_ACT(info) && !(st[u] & JB_PS_INFECTED);
      }
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) s0[u] = sus[u] ? a.susc[p[u]] : 0.0;
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const uint32_t sm = __ballot_sync(JB_FULL, sus[u]);
      if (sus[u]) {
        double lam_v[VM];
        double lam = 0.0;
#pragma unroll
        for (int v = 0; v < VM; ++v) {
          lam_v[v] = (v < V) ? row[v] * ((v == 0) ? s0[u] : a.susc[(size_t)v * a.NP + p[u]]) : 0.0;
          lam += lam_v[v];
        }
        if (a.lambda) a.lambda[p[u]] = lam;
        if (lam != 0.0) jb_draw_core<VM>(a, p[u], lam_v, lam, ord + __popc(sm & lane_lt));  // lambda == 0 never infects
      }
      ord += __popc(sm);
    }
  }
  if (lane == 0) atomicAdd(&a.counters[CNT_DRAWS], nsus);
}

// ---- tile kernel ----------------------------------------------------------------------------
// Small groups (<= 32 registered slots, contact matrix dim <= 4, static members) are packed at
// world-create time into 32-slot tiles that never split a group; one warp processes one tile
// entirely in registers.  STREAM: the tile's slots ARE the (renumbered) person indices, so the
// person state is read as a coalesced stream with no index indirection (households).  Otherwise
// the slot holds the person index and the state is gathered (companies, care homes, ...).
// Tiles without a present infector are dismissed after reading one state byte per slot.
#define JB_TI_VALID 0x0800u  // slot holds a person
#define JB_TI_FIRST 0x1000u  // first slot of its group

// Kernel A: streaming detection.  One thread owns 16 consecutive slots (a lane pair owns a
// tile): 128-bit loads of the slot descriptors and state bytes, per-tile bit masks exchanged with
// one shuffle, and for the (few) groups that hold both a present infector and a present
// susceptible one work item {first slot, length, group} is appended to the supergroup's work list
// (one atomic per warp iteration).  Everything else in the tile costs a handful of bit operations.
__device__ __forceinline__ uint32_t jb_byte(const uint4& v, int i) {
  const uint32_t w = i < 4 ? v.x : i < 8 ? v.y : i < 12 ? v.z : v.w;
  return (w >> (8 * (i & 3))) & 0xFFu;
}
__device__ __forceinline__ uint32_t jb_half(const uint4& v, int i) {
  const uint32_t w = i < 2 ? v.x : i < 4 ? v.y : i < 6 ? v.z : v.w;
  return (w >> (16 * (i & 1))) & 0xFFFFu;
}

template <bool STREAM>
__global__ void __launch_bounds__(256) k_detect(const __grid_constant__ GroupArgs a) {
  const int lane = threadIdx.x & 31;
  uint32_t* const my_cnt = a.work_cnt + (blockIdx.x & (JB_WORK_SUBLISTS - 1));
  uint2* const my_items = a.work_items + a.sub_off[blockIdx.x & (JB_WORK_SUBLISTS - 1)];
  const uint32_t n_chunks = a.n_tiles * 2;
  const uint32_t n_threads = gridDim.x * blockDim.x;
  const uint32_t n_iter = (n_chunks + n_threads - 1) / n_threads;
  uint32_t draws = 0;
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  for (uint32_t it = 0; it < n_iter; ++it, c += n_threads) {
    const bool ok = c < n_chunks;
    const uint32_t slot0 = c * 16;
    uint32_t valid = 0, first = 0, pres = 0, inf = 0;
    if (ok) {
      const uint4 i0 = *reinterpret_cast<const uint4*>(a.tp_info + slot0);
      const uint4 i1 = *reinterpret_cast<const uint4*>(a.tp_info + slot0 + 8);
      uint32_t sw[4];
      if (STREAM) {
        const uint4 sv = *reinterpret_cast<const uint4*>(a.pstate + a.stream_base + slot0);
        sw[0] = sv.x; sw[1] = sv.y; sw[2] = sv.z; sw[3] = sv.w;
      } else {
        uint32_t mem[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 mv = *reinterpret_cast<const uint4*>(a.tp_member + slot0 + 4 * q);
          mem[4 * q] = mv.x; mem[4 * q + 1] = mv.y; mem[4 * q + 2] = mv.z; mem[4 * q + 3] = mv.w;
        }
        uint32_t stg[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) stg[i] = mem[i] != 0xFFFFFFFFu ? a.pstate[mem[i]] : 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) sw[q] = stg[4 * q] | (stg[4 * q + 1] << 8) | (stg[4 * q + 2] << 16) | (stg[4 * q + 3] << 24);
      }
      // four slots at a time: the high byte of each descriptor holds activity (bits 0-2), VALID
      // (bit 3) and FIRST (bit 4); one byte-wise compare against the four state bytes
      const uint32_t iw[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t hi4 = __byte_perm(iw[2 * q], iw[2 * q + 1], 0x7531);
        const uint32_t st4 = sw[q];
        const uint32_t valid4 = (hi4 >> 3) & 0x01010101u;
        const uint32_t first4 = (hi4 >> 4) & valid4;
        const uint32_t pres4 = __vcmpeq4(hi4 & 0x07070707u, st4 & 0x07070707u) & valid4;
        const uint32_t inf4 = (st4 >> 3) & pres4;
        // bit 0 of every byte -> 4-bit mask
        valid |= (((valid4 * 0x00204081u) >> 21) & 0xFu) << (4 * q);
        first |= (((first4 * 0x00204081u) >> 21) & 0xFu) << (4 * q);
        pres |= (((pres4 * 0x00204081u) >> 21) & 0xFu) << (4 * q);
        inf |= (((inf4 * 0x00204081u) >> 21) & 0xFu) << (4 * q);
      }
    }
    // a tile = two neighbouring chunks: exchange the 16-bit masks inside the lane pair
    const uint32_t mine = first | (pres << 16);
    const uint32_t other = __shfl_xor_sync(JB_FULL, mine, 1);
    const uint32_t iv_o = __shfl_xor_sync(JB_FULL, inf | (valid << 16), 1);
    uint32_t items[4], item_g[4], n_items = 0, n_extra = 0;
    const uint32_t t = c >> 1;
    if (ok && !(lane & 1)) {  // the even lane of the pair owns the tile
      const uint32_t first32 = (first & 0xFFFFu) | (other << 16);
      const uint32_t pres32 = pres | (other & 0xFFFF0000u);
      const uint32_t inf32 = inf | (iv_o << 16);
      const uint32_t n_valid = 32 - __clz((valid & 0xFFFFu) | (iv_o & 0xFFFF0000u));  // valid slots are a prefix
      const uint32_t sus32 = pres32 & ~inf32;
      uint32_t rem = inf32;
      const uint32_t tg0 = rem ? a.tile_group0[t] : 0u;
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
        if (a.mode == 1) { a.draw_cnt[g] = nsus; continue; }
        draws += nsus;
        const uint32_t len = end - start;
        const uint32_t item = ((t * 32 + start) << 5) | (len - 1);
        if (n_items < 4) { items[n_items] = item; item_g[n_items] = g; ++n_items; }
        else {  // rare: more than 4 groups with infectors in one tile
          const uint32_t pos = atomicAdd(my_cnt, 1u);
          my_items[pos] = make_uint2(item, g);
          ++n_extra;
        }
      }
    }
    (void)n_extra;
    // one atomic per warp iteration for all buffered items
    uint32_t incl = n_items;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t up = __shfl_up_sync(JB_FULL, incl, off);
      if (lane >= off) incl += up;
    }
    const uint32_t total = __shfl_sync(JB_FULL, incl, 31);
    if (total) {
      uint32_t base = 0;
      if (lane == 31) base = atomicAdd(my_cnt, total);
      base = __shfl_sync(JB_FULL, base, 31) + incl - n_items;
      for (uint32_t q = 0; q < n_items; ++q) my_items[base + q] = make_uint2(items[q], item_g[q]);
    }
  }
  draws = __reduce_add_sync(JB_FULL, draws);
  if (lane == 0 && draws) atomicAdd(&a.counters[CNT_DRAWS], draws);
}

// Kernel B: 8 lanes per group that must timestep (4 groups per warp).  The lanes load the
// members' state in parallel (one dependent round per 8 members instead of one per member), then
// every lane accumulates the subgroup sizes and transmission sums from shuffle broadcasts in slot
// order -- the reference's own left-to-right order (Python `sum`, interaction.py:463-467) -- and
// finally each lane draws for its own member.
#define JB_DRAW_QUEUE 64  // per-warp queue of pending Bernoulli draws (shared memory)

template <bool STREAM, int VM, int JB_LG>
__global__ void __launch_bounds__(128) k_process(const __grid_constant__ GroupArgs a) {
  // Philox + exp are the expensive part and only ~1 lane in 3 holds a susceptible: the draws are
  // queued per warp and executed 32 at a time on fully populated warps
  __shared__ double q_lam[4][VM][JB_DRAW_QUEUE];
  __shared__ uint32_t q_p[4][JB_DRAW_QUEUE], q_ord[4][JB_DRAW_QUEUE];
  // lane j holds the length of sub-list j; exclusive prefix -> flat work index space
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t my_len = a.work_cnt[lane];
  uint32_t my_pre = my_len;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t up = __shfl_up_sync(JB_FULL, my_pre, off);
    if (lane >= off) my_pre += up;
  }
  const uint32_t n_work = __shfl_sync(JB_FULL, my_pre, 31);
  my_pre -= my_len;  // exclusive
  const uint32_t my_region = a.sub_off[lane];
  const int V = (VM > 1) ? a.V : 1;
  const int sl = lane & (JB_LG - 1);
  const uint32_t submask = ((1u << JB_LG) - 1u) << (lane & ~(JB_LG - 1));
  const uint32_t lane_lt = (1u << lane) - 1u;
  constexpr uint32_t per_warp = 32 / JB_LG;
  const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
  uint32_t qn = 0;  // warp-uniform queue length

  auto drain = [&](uint32_t n) {
    if ((uint32_t)lane < n) {
      double lam_v[VM];
      double lam = 0.0;
#pragma unroll
      for (int v = 0; v < VM; ++v) { lam_v[v] = q_lam[wib][v][lane]; lam += lam_v[v]; }
      jb_draw_core<VM>(a, q_p[wib][lane], lam_v, lam, q_ord[wib][lane]);
    }
    __syncwarp();
  };

  // all lanes of a warp iterate together (the queue is warp-wide); sub-groups beyond the work
  // list idle through the iteration
  const uint32_t warp0 = blockIdx.x * (blockDim.x >> 5) + wib;
  for (uint32_t wbase = warp0 * per_warp; wbase < n_work; wbase += n_warps * per_warp) {
    const uint32_t w = wbase + (lane / JB_LG);
    const bool have_item = w < n_work;
    // sub-list of flat index w: the largest j whose exclusive prefix is <= w (binary search over
    // the prefix held one entry per lane)
    uint2 item = make_uint2(0u, 0u);
    {
      uint32_t j = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const uint32_t cand = j + step;
        const uint32_t pre_c = __shfl_sync(JB_FULL, my_pre, cand & 31);
        if (cand < JB_WORK_SUBLISTS && pre_c <= w) j = cand;
      }
      const uint32_t pre_j = __shfl_sync(JB_FULL, my_pre, j), reg_j = __shfl_sync(JB_FULL, my_region, j);
      if (have_item) item = a.work_items[reg_j + (w - pre_j)];
    }
    const uint32_t slot0 = item.x >> 5, len = have_item ? (item.x & 31u) + 1u : 0u, g = item.y;
    int n[4] = {0, 0, 0, 0};
    double T[VM][4];
#pragma unroll
    for (int v = 0; v < VM; ++v)
#pragma unroll
      for (int k = 0; k < 4; ++k) T[v][k] = 0.0;
    // my own member of the first chunk stays in registers for the draw pass
    uint32_t p0 = 0, pk0 = 0;
    for (uint32_t base = 0; base < len; base += JB_LG) {
      const uint32_t m = base + sl;
      uint32_t pk = 0, p = 0;  // lsg | present << 2 | infected << 3 | variant << 4
      double tv = 0.0;
      if (m < len) {
        const uint32_t info = a.tp_info[slot0 + m];
        p = STREAM ? a.stream_base + slot0 + m : a.tp_member[slot0 + m];
        const uint32_t st = a.pstate[p];
        if ((st & JB_PS_ACT_MASK) == JB_INFO_ACT(info)) {
          pk = (JB_INFO_LSG(info) & 3u) | 4u;
          if (st & JB_PS_INFECTED) {
            pk |= 8u;
            tv = a.trans[p];
            if (VM > 1) pk |= (uint32_t)a.pvar[p] << 4;
          }
        }
      }
      if (base == 0) { p0 = p; pk0 = pk; }
      const uint32_t cnt = min((uint32_t)JB_LG, len - base);
      for (uint32_t k = 0; k < cnt; ++k) {
        const uint32_t q = __shfl_sync(submask, pk, k, JB_LG);
        const double t = __shfl_sync(submask, tv, k, JB_LG);
        if (q & 4u) {
          const int l = q & 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) n[j] += (j == l);
          if (q & 8u) {
            const int pv = (VM > 1) ? (int)(q >> 4) : 0;
#pragma unroll
            for (int v = 0; v < VM; ++v)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (v == pv && j == l) T[v][j] += t;
          }
        }
      }
    }
    double beta = 0.0;
    uint32_t ord = 0;
    if (have_item) {
      const uint32_t gi = a.grp_info[g];
      beta = a.beta[JB_GINFO_BETA(gi) * a.R + JB_GINFO_REGION(gi)] * a.beta_scale[JB_GINFO_SCALE(gi)];
      ord = a.inject ? a.draw_base[g] : 0u;
    }
    const uint32_t max_len = __reduce_max_sync(JB_FULL, len);
    for (uint32_t base = 0; base < max_len; base += JB_LG) {
      const uint32_t m = base + sl;
      uint32_t pk = pk0, p = p0;
      if (base != 0) {
        pk = 0;
        if (m < len) {
          const uint32_t info = a.tp_info[slot0 + m];
          p = STREAM ? a.stream_base + slot0 + m : a.tp_member[slot0 + m];
          const uint32_t st = a.pstate[p];
          if ((st & JB_PS_ACT_MASK) == JB_INFO_ACT(info)) pk = (JB_INFO_LSG(info) & 3u) | 4u | ((st & JB_PS_INFECTED) ? 8u : 0u);
        }
      }
      const bool sus = m < len && (pk & 12u) == 4u;
      const uint32_t susm = __ballot_sync(JB_FULL, sus);
      const uint32_t sub_sus = (susm >> (lane & ~(JB_LG - 1))) & ((1u << JB_LG) - 1u);
      double lam_v[VM];
      double lam = 0.0;
      if (sus) {
        const int i = pk & 3;
#pragma unroll
        for (int v = 0; v < VM; ++v) {
          double r = 0.0;
          if (v < V) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < a.dim && T[v][j] != 0.0) {
                const int nn = (i == j) ? max(1, n[j] - 1) : n[j];  // interaction.py:469-471
                r += a.cm[i * a.dim + j] * T[v][j] / (double)nn * beta * a.sp->dt;
              }
            r *= a.susc[(size_t)v * a.NP + p];
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
          q_p[wib][pos] = p;
          q_ord[wib][pos] = ord + __popc(sub_sus & ((1u << sl) - 1u));
#pragma unroll
          for (int v = 0; v < VM; ++v) q_lam[wib][v][pos] = lam_v[v];
        }
        qn += __popc(wm);
        __syncwarp();
        if (qn >= 32) {
          drain(32);
          const uint32_t rest = qn - 32;
          uint32_t mp = 0, mo = 0;
          double ml[VM];
          if ((uint32_t)lane < rest) {
            mp = q_p[wib][32 + lane]; mo = q_ord[wib][32 + lane];
#pragma unroll
            for (int v = 0; v < VM; ++v) ml[v] = q_lam[wib][v][32 + lane];
          }
          __syncwarp();
          if ((uint32_t)lane < rest) {
            q_p[wib][lane] = mp; q_ord[wib][lane] = mo;
#pragma unroll
            for (int v = 0; v < VM; ++v) q_lam[wib][v][lane] = ml[v];
          }
          __syncwarp();
          qn = rest;
        }
      }
      ord += __popc(sub_sus);
    }
  }
  if (qn) drain(qn);
}

// ------------------------------------------------------------------------------------------
// Stage C1: create infections (InfectionSelector._make_infection).
// ------------------------------------------------------------------------------------------
struct NewInfArgs {
  const uint32_t* member;   // person indices (local, or >= N for halo: skipped here)
  const uint8_t* variant;
  const uint32_t* count_dev;  // number of entries (device) or null -> count_host
  uint32_t count_host;
  uint32_t cap;
  uint32_t N, NP, C;
  int V;
  uint32_t gid_base;
  const uint32_t* int2ext;
  const StepDev* sp;        // now, seed, step, flags
  const uint8_t* age; const uint8_t* sex; const uint8_t* ch;
  const double* hi_cum;
  const DiseaseDev* dis;
  uint8_t* pstate; uint8_t* pvar; int8_t* ptag; double* susc; double* trans; int32_t* pslot;
  int32_t* inf_person; double* inf_start; double* inf_par; double* inf_next; double* traj_time;
  uint8_t* inf_stage; uint8_t* inf_nst; uint8_t* inf_var; int8_t* inf_tag; int8_t* inf_maxtag; int8_t* traj_tag;
  uint8_t* inf_region; int32_t* inf_hosp; int32_t* inf_med; uint8_t* inf_sev;  // slot-resident person facts
  const uint32_t* psa; const int32_t* sa_hospital; const int32_t* pmed;
  uint32_t* counters;
  uint32_t* summary;        // [R][JB_N_SUMMARY]
  const uint8_t* pregion;
  uint8_t* halo_ret;        // per halo person: variant+1 if infected here (or null)
};

// 16 lanes per new infection: every completion time / transmission parameter is an independent
// sample with its own Philox stream (JB_STREAM_SAMPLE0 + slot), so the lanes draw them in
// parallel and lane 0 assembles the infection.  Slots: 0 max_severity; 1..8 stage s = slot-1;
// 9..14 transmission parameter k = slot-9; 15 the second `alpha()` call of the xnexp profile.
__global__ void __launch_bounds__(128) k_newinf(const __grid_constant__ NewInfArgs a) {
  // the disease tables are walked with data-dependent indices: stage them in shared memory once
  // per block instead of chasing them through L2 for every sample
  __shared__ __align__(8) unsigned char s_dis[sizeof(DiseaseDev)];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.dis);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_dis);
    for (uint32_t i = threadIdx.x; i < sizeof(DiseaseDev) / 4; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const uint32_t n = a.count_dev ? min(*a.count_dev, a.cap) : a.count_host;
  const DiseaseDev& D = *reinterpret_cast<const DiseaseDev*>(s_dis);
  const int lane = threadIdx.x & 31, j = lane & 15;
  const uint32_t segmask = 0xFFFFu << (lane & 16);
  const uint32_t n_seg = gridDim.x * blockDim.x / 16;
  for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) / 16; i < n; i += n_seg) {
    const uint32_t p = a.member[i];
    const uint32_t var = a.variant ? a.variant[i] : 0u;
    if (p >= a.N) {  // halo person: tell the home domain (epidemiology.py:359-401)
      if (j == 0) {
        if (a.halo_ret) a.halo_ret[p - a.N] = (uint8_t)(var + 1);
        atomicAdd(&a.counters[CNT_HALO_INF], 1u);
      }
      continue;
    }
    if (a.sp->flags & JB_STEP_NO_NEW_INFECTIONS) continue;  // only route halo infections (the host creates the local ones)
    const uint32_t st = a.pstate[p];
    if (st & (JB_PS_INFECTED | JB_PS_DEAD)) continue;  // cannot happen within a step; guards seeding
    const uint32_t gid = a.gid_base + a.int2ext[p];
    // symptoms: max_severity -> outcome slot (symptoms.py:47,110-120; health_index.py:151-182)
    JbStream r0;
    jb_stream_init(r0, a.sp->seed, a.sp->step, gid, JB_STREAM_SAMPLE0);
    const double max_sev = jb_stream_next(r0);
    const int age = a.age[p] > 99 ? 99 : a.age[p];
    const int pop = (a.ch[p] && age >= 50) ? 1 : 0;
    const double* cum = a.hi_cum + ((size_t)(pop * 2 + a.sex[p]) * 100 + age) * D.n_tags;
    int idx = 0;
    while (idx < D.n_tags && cum[idx] < max_sev) ++idx;  // np.searchsorted(side="left")
    if (idx >= D.n_tags) idx = D.n_tags - 1;  // residual mass of a cumsum ending at 1-eps
    int k = 0;
    for (int t = 0; t < D.n_traj; ++t)
      if (D.traj_max_tag[t] == idx) k = t;
    const int nst = D.traj_n_stages[k];
    // my sample
    double x = 0.0;
    {
      JbStream rs;
      jb_stream_init(rs, a.sp->seed, a.sp->step, gid, JB_STREAM_SAMPLE0 + (uint32_t)j);
      if (j >= 1 && j <= 8) {
        const int s = j - 1;
        if (s < nst) x = jb_sample_dist(rs, D.traj_stage_time[k][s].type, D.traj_stage_time[k][s].p);
      } else if (j >= 9) {
        const int q = (j == 15) ? 3 : j - 9;
        const int need = D.trans_type == JB_TRANS_GAMMA ? 4 : D.trans_type == JB_TRANS_XNEXP ? 5 : 1;
        if (q < need && (j != 15 || D.trans_type == JB_TRANS_XNEXP)) x = jb_sample_dist(rs, D.trans_par[q].type, D.trans_par[q].p);
      }
    }
    // lane 0 of the segment assembles the infection
    double tstage[JB_MAX_STAGES], tp[7];
#pragma unroll
    for (int s = 0; s < JB_MAX_STAGES; ++s) tstage[s] = __shfl_sync(segmask, x, 1 + s, 16);
#pragma unroll
    for (int q = 0; q < 7; ++q) tp[q] = __shfl_sync(segmask, x, 9 + q, 16);
    if (j != 0) continue;
    const uint32_t slot = atomicAdd(&a.counters[CNT_SLOTS_HW], 1u);
    if (slot >= a.C) { atomicAdd(&a.counters[CNT_OVERFLOW], 1u); continue; }
    // trajectory: cumulative stage start times (trajectory_maker.py:215-227)
    double cumt = 0.0, t_exposed = 0.0;
#pragma unroll
    for (int s = 0; s < JB_MAX_STAGES; ++s) {
      double tt = 1e300;
      int tg = 0;
      if (s < nst) {
        tt = cumt;
        tg = D.traj_stage_tag[k][s];
        cumt += tstage[s];
      }
      if (s == 1) t_exposed = tt;
      a.traj_time[(size_t)s * a.C + slot] = tt;
      a.traj_tag[(size_t)s * a.C + slot] = (int8_t)tg;
    }
    // transmission (infection_selector.py:280-332)
    double par[JB_N_TPAR] = {0, 0, 0, 0, 0};
    if (D.trans_type == JB_TRANS_GAMMA) {
      par[0] = tp[1]; par[1] = tp[3] + t_exposed; par[2] = 1.0 / tp[2]; par[3] = tp[0];
    } else if (D.trans_type == JB_TRANS_XNEXP) {
      const double tfi = tp[1] + t_exposed;
      const double peak = t_exposed - tfi + tp[2];
      const double norm_time = tp[4];
      const double nn = peak / tp[3];
      const double alpha = tp[6];
      const double max_tau = nn * alpha * norm_time / norm_time;  // transmission_xnexp.py:119-121
      par[0] = nn; par[1] = alpha; par[2] = tfi; par[4] = norm_time;
      par[3] = tp[0] / (pow(max_tau, nn) * exp(-max_tau / alpha));
    } else {
      par[0] = tp[0];
    }
    for (int q = 0; q < JB_N_TPAR; ++q) a.inf_par[(size_t)q * a.C + slot] = par[q];
    a.inf_person[slot] = (int32_t)p;
    a.inf_start[slot] = a.sp->now;
    a.inf_next[slot] = nst > 1 ? tstage[0] : 1e300;  // == traj_time[1]
    a.inf_stage[slot] = 0;
    a.inf_nst[slot] = (uint8_t)nst;
    a.inf_var[slot] = (uint8_t)var;
    a.inf_tag[slot] = (int8_t)D.tag_exposed;
    a.inf_maxtag[slot] = (int8_t)idx;
    a.inf_region[slot] = a.pregion[p];
    a.inf_hosp[slot] = a.sa_hospital[a.psa[p]];
    a.inf_med[slot] = a.pmed[p];
    a.inf_sev[slot] = 0;
    a.pslot[p] = (int32_t)slot;
    a.pvar[p] = (uint8_t)var;
    a.ptag[p] = (int8_t)D.tag_exposed;
    a.trans[p] = (D.trans_type == JB_TRANS_CONSTANT) ? par[0] : 0.0;
    // immunity.add_immunity(infection.immunity_ids()) (infection_selector.py:160-162)
    for (int v = 0; v < a.V; ++v) a.susc[(size_t)v * a.NP + p] = 0.0;
    a.pstate[p] = (uint8_t)(st | JB_PS_INFECTED);
    atomicAdd(&a.summary[a.pregion[p] * JB_N_SUMMARY + SUM_NEW_INFECTED], 1u);
  }
}

// ------------------------------------------------------------------------------------------
// Stage C2: health update for every active infection slot (epidemiology.py:217-303).
// ------------------------------------------------------------------------------------------
struct DiseaseMasks {
  int trans_type, rec_mask, dead_mask, hosp_mask, icu_mask, severe_mask;
};

struct HealthArgs {
  DiseaseMasks masks;
  uint32_t C;
  const StepDev* sp;  // t_now = timer.now + timer.duration, flags
  const DiseaseDev* dis;
  int32_t* inf_person; const double* inf_start; const double* inf_par; double* inf_next; const double* traj_time;
  uint8_t* inf_stage; const uint8_t* inf_nst; int8_t* inf_tag; const int8_t* traj_tag;
  const uint8_t* inf_region; const int32_t* inf_hosp; int32_t* inf_med; uint8_t* inf_sev;
  uint8_t* pstate; int8_t* ptag; double* trans; int32_t* pmed; int32_t* pslot; uint8_t* phas;
  uint32_t* counters;
  uint32_t* summary;        // [R][JB_N_SUMMARY]
  int R;
};

#define JB_SUM_SMEM_REGIONS 64

__device__ __forceinline__ double jb_gamma_pdf(double x, double a, double loc, double scale) {
  // transmission.py:47-75
  if (x < loc) return 0.0;
  double z = (x - loc) / scale;
  return 1.0 / tgamma(a) * pow(z, a - 1.0) * exp(-z) / scale;
}

// One thread per infection slot.  Everything the update needs about the PERSON (region, closest
// hospital, ward, severe flag) is mirrored in the slot, so the only per-step access indexed by
// person is the write of the new transmission probability; the person's state byte, ward and
// tag are touched only when they change.
__global__ void __launch_bounds__(256) k_health(const __grid_constant__ HealthArgs a) {
  // per-region counters are accumulated in shared memory and flushed once per block: a global
  // atomic per infected person on nine addresses serialises in L2
  __shared__ uint32_t s_sum[JB_SUM_SMEM_REGIONS * JB_N_SUMMARY];
  const bool sum_smem = a.R <= JB_SUM_SMEM_REGIONS;
  if (sum_smem) {
    for (int i = threadIdx.x; i < a.R * JB_N_SUMMARY; i += blockDim.x) s_sum[i] = 0;
    __syncthreads();
  }
  const uint32_t hw = min(a.counters[CNT_SLOTS_HW], a.C);
  const DiseaseMasks D = a.masks;
  const double t_now = a.sp->t_now;
  const uint32_t flags = a.sp->flags;
  uint32_t c_inf = 0, c_hosp = 0, c_icu = 0;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < hw; s += gridDim.x * blockDim.x) {
    const int32_t p = a.inf_person[s];
    if (p < 0) continue;
    const double t = t_now - a.inf_start[s];
    // transmission.update_infection_probability (infection.py:91)
    if (D.trans_type == JB_TRANS_GAMMA) {
      a.trans[p] = a.inf_par[(size_t)3 * a.C + s] *
                   jb_gamma_pdf(t, a.inf_par[s], a.inf_par[(size_t)1 * a.C + s], a.inf_par[(size_t)2 * a.C + s]);
    } else if (D.trans_type == JB_TRANS_XNEXP) {
      const double tfi = a.inf_par[(size_t)2 * a.C + s];
      double prob = 0.0;
      if (t > tfi) {
        const double tau = (t - tfi) / a.inf_par[(size_t)4 * a.C + s];
        prob = a.inf_par[(size_t)3 * a.C + s] * (pow(tau, a.inf_par[s]) * exp(-tau / a.inf_par[(size_t)1 * a.C + s]));
      }
      a.trans[p] = prob;
    }
    // symptoms.update_trajectory_stage: at most one stage per call (symptoms.py:152-154)
    int stage = a.inf_stage[s];
    int tag = a.inf_tag[s];
    const int nst = a.inf_nst[s];
    bool changed = false;
    if (stage + 1 < nst && t > a.inf_next[s]) {
      ++stage;
      tag = a.traj_tag[(size_t)stage * a.C + s];
      a.inf_stage[s] = (uint8_t)stage;
      a.inf_tag[s] = (int8_t)tag;
      a.inf_next[s] = (stage + 1 < JB_MAX_STAGES) ? a.traj_time[(size_t)(stage + 1) * a.C + s] : 1e300;
      changed = true;
    }
    const int bit = JB_TAG_BIT(tag);
    uint32_t* sum = (sum_smem ? s_sum : a.summary) + (size_t)a.inf_region[s] * JB_N_SUMMARY;
    // Hospitalisation.apply (medical_care_policies.py:413-526; hospital.py:62-122)
    const int32_t med_old = a.inf_med[s];
    const bool was_med = med_old >= 0;
    int32_t med_new = med_old;
    if (flags & JB_STEP_HOSPITALISATION) {
      if (bit & D.hosp_mask) {
        const int32_t hg = was_med ? (med_old >> 8) : a.inf_hosp[s];
        if (hg >= 0) med_new = (hg << 8) | ((bit & D.icu_mask) ? 2 : 1);
      } else if (was_med && !(bit & D.dead_mask)) {
        med_new = -1;
      }
    }
    const uint32_t sev_new = (bit & D.severe_mask) ? 1u : 0u;
    const bool rec = (bit & D.rec_mask) != 0, dead = !rec && (bit & D.dead_mask) != 0;
    if (rec || dead || changed || med_new != med_old || sev_new != a.inf_sev[s]) {
      uint32_t st = a.pstate[p];
      st = (med_new >= 0) ? (st | JB_PS_MEDICAL) : (st & ~JB_PS_MEDICAL);
      st = sev_new ? (st | JB_PS_SEVERE) : (st & ~JB_PS_SEVERE);
      if (med_new != med_old) { a.pmed[p] = med_new; a.inf_med[s] = med_new; }
      a.inf_sev[s] = (uint8_t)sev_new;
      if (rec) {  // Epidemiology.recover (epidemiology.py:197-215)
        st &= ~(JB_PS_INFECTED | JB_PS_SEVERE);
        a.inf_person[s] = -1;
        a.pslot[p] = -1;
        a.trans[p] = 0.0;
      } else if (dead) {  // bury_the_dead (epidemiology.py:168-195)
        st &= ~(JB_PS_INFECTED | JB_PS_SEVERE | JB_PS_MEDICAL);
        st |= JB_PS_DEAD;
        a.inf_person[s] = -1;
        a.pslot[p] = -1;
        a.pmed[p] = -1;
        a.phas[p] = 0;
        a.trans[p] = 0.0;
      }
      // ptag mirrors person.infection.tag; -1 ("healthy") once the infection object is gone
      a.ptag[p] = (rec || dead) ? (int8_t)-1 : (int8_t)tag;
      a.pstate[p] = (uint8_t)st;
    }
    if (rec) {
      atomicAdd(&a.counters[CNT_RECOVERED_STEP], 1u);
      atomicAdd(&sum[SUM_RECOVERED], 1u);
    } else if (dead) {
      atomicAdd(&a.counters[CNT_DEATHS_STEP], 1u);
      atomicAdd(&a.counters[CNT_DEAD_TOTAL], 1u);
      atomicAdd(&sum[SUM_DEATHS], 1u);
      if (was_med) atomicAdd(&sum[SUM_HOSP_DEATHS], 1u);
    } else {
      ++c_inf;
      atomicAdd(&sum[SUM_CUR_INFECTED], 1u);
      if (med_new >= 0) {
        ++c_hosp;
        atomicAdd(&sum[SUM_CUR_HOSP], 1u);
        if (bit & D.icu_mask) { ++c_icu; atomicAdd(&sum[SUM_CUR_ICU], 1u); }
        if (!was_med) atomicAdd(&sum[SUM_ADMISSIONS], 1u);
      }
    }
  }
  // block-level reduction of the "current" counters
  c_inf = __reduce_add_sync(JB_FULL, c_inf);
  c_hosp = __reduce_add_sync(JB_FULL, c_hosp);
  c_icu = __reduce_add_sync(JB_FULL, c_icu);
  if ((threadIdx.x & 31) == 0) {
    if (c_inf) atomicAdd(&a.counters[CNT_INFECTED], c_inf);
    if (c_hosp) atomicAdd(&a.counters[CNT_HOSP], c_hosp);
    if (c_icu) atomicAdd(&a.counters[CNT_ICU], c_icu);
  }
  if (sum_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < a.R * JB_N_SUMMARY; i += blockDim.x)
      if (s_sum[i]) atomicAdd(&a.summary[i], s_sum[i]);
  }
}

// ------------------------------------------------------------------------------------------
// Halo exchange kernels (the device form of MovablePeople, mpi_wrapper.py:145-293).
// Record: 8 bytes header {pstate, variant, 6 pad} + max(1,V) doubles
// (infected: [0] = transmission probability; else: susceptibility per variant).
// ------------------------------------------------------------------------------------------
__global__ void k_halo_pack(const uint32_t* __restrict__ out_person, uint32_t n_out, const uint8_t* __restrict__ pstate,
                            const uint8_t* __restrict__ pvar, const double* __restrict__ susc,
                            const double* __restrict__ trans, uint32_t NP, int V, unsigned char* __restrict__ send) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  uint32_t p = out_person[i];
  size_t rec = 8 + (size_t)8 * V;
  unsigned char* r = send + (size_t)i * rec;
  uint32_t st = pstate[p];
  uint64_t hdr = (uint64_t)st | ((uint64_t)pvar[p] << 8);
  *(uint64_t*)r = hdr;
  double* d = (double*)(r + 8);
  if (st & JB_PS_INFECTED) {
    d[0] = trans[p];
    for (int v = 1; v < V; ++v) d[v] = 0.0;
  } else {
    for (int v = 0; v < V; ++v) d[v] = susc[(size_t)v * NP + p];
  }
}

__global__ void k_halo_unpack(const unsigned char* __restrict__ recv, uint32_t H, uint32_t N, uint32_t NP, int V,
                              uint8_t* __restrict__ pstate, uint8_t* __restrict__ pvar, double* __restrict__ susc,
                              double* __restrict__ trans, uint8_t* __restrict__ halo_ret) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H) return;
  size_t rec = 8 + (size_t)8 * V;
  const unsigned char* r = recv + (size_t)j * rec;
  uint64_t hdr = *(const uint64_t*)r;
  const double* d = (const double*)(r + 8);
  uint32_t p = N + j;
  uint32_t st = (uint32_t)(hdr & 0xFFu);
  pstate[p] = (uint8_t)st;
  pvar[p] = (uint8_t)((hdr >> 8) & 0xFFu);
  if (st & JB_PS_INFECTED) {
    trans[p] = d[0];
  } else {
    for (int v = 0; v < V; ++v) susc[(size_t)v * NP + p] = d[v];
  }
  if (halo_ret) halo_ret[j] = 0;
}

// Returned bytes (variant+1 per outbound slot) -> list of residents to infect.
__global__ void k_halo_return(const uint8_t* __restrict__ ret, const uint32_t* __restrict__ out_person, uint32_t n_out,
                              uint32_t* __restrict__ member, uint8_t* __restrict__ variant, uint32_t cap,
                              uint32_t* __restrict__ counter) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  uint32_t v = ret[i];
  if (!v) return;
  uint32_t k = atomicAdd(counter, 1u);
  if (k < cap) { member[k] = out_person[i]; variant[k] = (uint8_t)(v - 1); }
}

__global__ void k_fill_f64(double* x, size_t n, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}
