// Hand-written sm_100a kernels of the JUNE per-timestep hot path.
//
//   k_place          ActivityManager.move_people_to_active_subgroups   (activity_manager.py:255-402)
//   k_tiles          Interaction.time_step_for_group, groups of <= 32 slots (jb_groups.cuh)
//   k_groups_cta     same, one CTA per group (schools, hospitals, big companies, ...)
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

// Transmission probabilities are double buffered ([2][NP] in one allocation): the interaction kernels of a
// step read half `tpar` (filled by the previous step's health update) while this step's health update fills
// the other half.  The half is a per-step scalar in the parameter block, not a kernel argument, so one
// captured graph serves both parities.
__device__ __forceinline__ void jb_prefetch_l2(const void* ptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}

__device__ __forceinline__ const double* jb_trans_read(const double* base, uint32_t NP, const StepDev* sp) {
  return base + (size_t)(sp->tpar & 1u) * NP;
}
__device__ __forceinline__ double* jb_trans_write(double* base, uint32_t NP, const StepDev* sp) {
  return base + (size_t)((sp->tpar & 1u) ^ 1u) * NP;
}

// ------------------------------------------------------------------------------------------
// Stage A: placement.  The reference walks the activity hierarchy per person and takes the first
// activity with a non-None subgroup (activity_manager.py:380-398); stay-home policies reduce the
// activity list to (medical_facility, residence) (individual_policies.py:87-119).  Here the
// chosen activity is written into the low bits of the person's per-step byte `phot` (together
// with the infected flag, susceptibility code and variant the interaction kernels need);
// fixed-membership subgroups test that byte instead of being scattered into.  People whose subgroup is dynamic
// (hospital patients) are scattered into the per-group bin of their ward.
// 16 persons per thread (one 128-bit load of each byte array), grid-stride, counters reduced in
// registers -> shared memory -> one global atomic per block and counter.
// ------------------------------------------------------------------------------------------
struct PlaceArgs {
  const uint8_t* pstate;
  uint8_t* phot;
  const uint8_t* pvar;   // null unless the world has more than one variant
  const uint8_t* phas;
  const int32_t* pmed;
  uint32_t* dyn_cnt;   // per dynamic group: members appended this step
  uint32_t* dyn_bin;   // bins of (lsg << 28) | person entries
  DynRanges dr;
  uint32_t* counters;
  uint32_t N;
  const StepDev* sp;
  // leisure
  LeisureDev lz;
  const uint8_t* age; const uint8_t* sex; const uint8_t* ch; const uint8_t* pregion;
  const uint32_t* int2ext;
  uint32_t gid_base;
  int R;
  CommuteDev cmt;
  // individual policies (null / 0 = none)
  const PolicyDev* pol;
  uint32_t n_pol;
  const int8_t* ptag;
  const uint8_t* pattr;
  const uint8_t* pwork_region;
  const double* pol_inject;  // [JB_MAX_POLICIES][JB_POLICY_DRAWS][N] or null
  int severe_mask;
  // mirror of the primary-activity supergroups (all null when the step has no mirrored supergroup to run):
  // table-driven placement: `pmir` = byte last pushed per person, changed bytes go to `mstate[pmslot[p]]`;
  // per-person placement: every mirrored person's byte of THIS step goes to `mstep[pmslot[p]]`
  uint8_t* mstate; uint8_t* mstep; uint8_t* pmir; const uint32_t* pmslot;
  const uint32_t* grp_info;  // region of a hospital
  uint32_t* summary;         // [R][JB_N_SUMMARY]: ward / intensive-care patients placed, by the hospital's region
  VisitsDev vz;              // residence visits (activity < 0: none)
  int vis_defer;             // this step resolves visits: leisure destinations are recorded, not scattered
};

// Uniform number j of policy k for person p in this step (or the injected one, test mode).
__device__ __forceinline__ double jb_policy_uniform(const PlaceArgs& a, uint32_t p, uint32_t k, uint32_t j) {
  if (a.pol_inject) return a.pol_inject[((size_t)k * JB_POLICY_DRAWS + j) * a.N + p];
  const JbPhilox4 r = jb_philox4x32_10(a.gid_base + a.int2ext[p], a.sp->step, JB_STREAM_POLICY0 + k, j >> 1,
                                       (uint32_t)a.sp->seed, (uint32_t)(a.sp->seed >> 32));
  return (j & 1u) ? jb_u52(r.x[2], r.x[3]) : jb_u52(r.x[0], r.x[1]);
}

// The first two uniforms of policy k for person p (one Philox block), or the injected ones.
__device__ __forceinline__ void jb_policy_uniform01(const PlaceArgs& a, uint32_t p, uint32_t k, double& u0, double& u1) {
  if (a.pol_inject) {
    u0 = a.pol_inject[((size_t)k * JB_POLICY_DRAWS + 0) * a.N + p];
    u1 = a.pol_inject[((size_t)k * JB_POLICY_DRAWS + 1) * a.N + p];
    return;
  }
  const JbPhilox4 r = jb_philox4x32_10(a.gid_base + a.int2ext[p], a.sp->step, JB_STREAM_POLICY0 + k, 0u,
                                       (uint32_t)a.sp->seed, (uint32_t)(a.sp->seed >> 32));
  u0 = jb_u52(r.x[0], r.x[1]);
  u1 = jb_u52(r.x[2], r.x[3]);
}

// CloseCompanies.check_skips_activity (individual_policies.py:651-770).  NaN = None.  The draws are consumed in order
// (u0, u1, then a third one); the first two come from ONE Philox block computed up front, so the branches of the
// decision tree hold no generator code (the lanes of a warp take different branches).
__device__ bool jb_close_companies_skips(const PlaceArgs& a, uint32_t p, uint32_t k, const JbIndividualPolicy& q, int status) {
  const double awp = q.p[0], fp = q.p[1], kp = q.p[2], fr = q.p[3], kr = q.p[4], rr = q.p[5];
  const bool have_f = !isnan(fr) && !isnan(fp), have_k = !isnan(kr) && !isnan(kp), have_r = !isnan(rr);
  if (q.flags & 1u) return true;  // full_closure
  double u[2];
  jb_policy_uniform01(a, p, k, u[0], u[1]);
  uint32_t j = 0;
  bool decided = false, skip = false;
  if (status == 3) {  // furlough
    if (!have_f) return true;
    if (fr < fp) return true;
    if (u[j++] < fp / fr) return true;
    return !isnan(awp) && u[j++] < awp;
  }
  if (status == 1 && have_k)  // key worker
    return kr > kp && u[j++] > kp / kr;
  if (status == 2 && !isnan(awp)) {  // random
    if (have_f && have_k && have_r) {
      if (fr < fp && kr < kp) {
        if (u[j++] < (fp - fr) / rr) { decided = true; skip = true; }
        else if (u[j++] < (kp - kr) / (rr - (fp - fr))) { decided = true; skip = false; }
      } else if (fr < fp) {
        if (u[j++] < (fp - fr) / rr) { decided = true; skip = true; }
      } else if (kr < kp) {
        if (u[j++] < (kp - kr) / rr) { decided = true; skip = false; }
      }
    } else if (have_f && have_r) {
      if (fr < fp && u[j++] < (fp - fr) / rr) { decided = true; skip = true; }
    } else if (have_k && have_r) {
      if (kr < kp && u[j++] < (kp - kr) / rr) { decided = true; skip = false; }
    }
    if (decided) return skip;
    const double last = j < 2 ? u[j] : jb_policy_uniform(a, p, k, j);  // (the third draw: both tests above failed)
    return last < awp;
  }
  return false;
}

// IndividualPolicies.apply (individual_policies.py:36-84): the person's allowed-activity mask.
__device__ __noinline__ uint32_t jb_apply_policies(const PlaceArgs& a, uint32_t p, uint32_t st, uint32_t allowed) {
  const uint32_t stay = (1u << JB_ACT_MEDICAL) | (1u << JB_ACT_RESIDENCE);
  const PolicyDev& P = *a.pol;
  const double rc = P.rc[a.pregion[p]];
  const uint32_t attr = a.pattr ? a.pattr[p] : 0u, kind = JB_PA_KIND(attr);
  const uint32_t age = a.age[p];
  const bool infected = (st & JB_PS_INFECTED) != 0;
  const int tagbit = infected ? JB_TAG_BIT(a.ptag[p]) : 0;
  // skip-activity policies only drop primary_activity / commute: when the step offers neither they are not evaluated
  // (their draws are keyed per policy, nothing else depends on them)
  const bool can_skip = (allowed & ((1u << JB_ACT_PRIMARY) | (1u << JB_ACT_COMMUTE))) != 0u;
  for (uint32_t k = 0; k < a.n_pol; ++k) {
    const JbIndividualPolicy& q = P.pol[k];
    if (!can_skip && q.type != JB_POL_SEVERE_STAY_HOME && q.type != JB_POL_QUARANTINE && q.type != JB_POL_SHIELDING) continue;
    switch (q.type) {
      case JB_POL_SEVERE_STAY_HOME:
        if (tagbit & a.severe_mask) return allowed & stay;
        break;
      case JB_POL_QUARANTINE: {
        // time_of_symptoms_onset is 0 for every infection (symptoms.py:77-88 stops at the first
        // trajectory entry), so release_day = n_days; the household branch never fires because a
        // quarantine_starting_date of 0 is falsy (household.py:175-185)
        const double release = 0.0 + q.p[0];
        if ((tagbit & (int)q.mask) && 0 < release - a.sp->now && release - a.sp->now < q.p[0])
          if (jb_policy_uniform(a, p, k, 0) < q.p[1] * rc) return allowed & stay;
        break;
      }
      case JB_POL_SHIELDING:
        if ((double)age >= q.p[0] && (isnan(q.p[1]) || jb_policy_uniform(a, p, k, 0) < q.p[1] * rc)) return allowed & stay;
        break;
      case JB_POL_CLOSE_SCHOOLS:
        if (kind == 1u) {
          bool skip = false;
          if (q.flags & 1u) skip = true;
          else if (!(age < 14u && (attr & JB_PA_TWO_KEY_WORKERS))) {
            if (age < 32u && ((q.mask >> age) & 1u)) skip = true;
            else if (jb_policy_uniform(a, p, k, 0) > q.p[0]) skip = true;
          }
          if (skip) allowed &= ~(1u << JB_ACT_PRIMARY);
        }
        break;
      case JB_POL_CLOSE_UNIVERSITIES:
        if (kind == 3u) allowed &= ~(1u << JB_ACT_PRIMARY);
        break;
      case JB_POL_CLOSE_COMPANIES:
        if (kind == 2u && jb_close_companies_skips(a, p, k, q, (int)JB_PA_LOCKDOWN(attr)))
          allowed &= ~((1u << JB_ACT_PRIMARY) | (1u << JB_ACT_COMMUTE));
        break;
      case JB_POL_CLOSE_COMPANIES_TIERS: {  // :556-601: one draw, whichever of the two conditions holds
        const uint32_t wr = a.pwork_region ? a.pwork_region[p] : 0xFFu;
        if (kind == 2u && JB_PA_LOCKDOWN(attr) == 2u && wr != 0xFFu) {
          const uint32_t home_t = (q.mask >> a.pregion[p]) & 1u, work_t = (q.mask >> wr) & 1u;
          if ((work_t || home_t) && jb_policy_uniform(a, p, k, 0) < rc)
            allowed &= ~((1u << JB_ACT_PRIMARY) | (1u << JB_ACT_COMMUTE));
        }
        break;
      }
      case JB_POL_LIMIT_LONG_COMMUTE:  // :815-824: skips when random() < going_to_work_probability
        if ((attr & JB_PA_LONG_COMMUTER) && jb_policy_uniform(a, p, k, 0) < q.p[0])
          allowed &= ~((1u << JB_ACT_PRIMARY) | (1u << JB_ACT_COMMUTE));
        break;
      default: break;
    }
  }
  return allowed;
}

// Append person p to the bin of dynamic group g (hospital ward, leisure venue).
__device__ __forceinline__ void jb_dyn_scatter(const PlaceArgs& a, uint32_t g, uint32_t lsg, uint32_t p) {
  for (uint32_t r = 0; r < a.dr.n; ++r) {
    const uint32_t k = g - a.dr.first[r];
    if (k < a.dr.count[r]) {
      const uint32_t i = atomicAdd(&a.dyn_cnt[a.dr.cnt_base[r] + k], 1u);
      // the bin is sorted later: key on the CALLER's person index, the order the reference appends in
      if (i < a.dr.cap[r]) { a.dyn_bin[(size_t)a.dr.slot_base[r] + (size_t)k * a.dr.cap[r] + i] = (lsg << 28) | a.int2ext[p]; return; }
      break;
    }
  }
  atomicAdd(&a.counters[CNT_OVERFLOW], 1u);
}

// A patient goes into its ward's bin; len(hospital.ward) / len(hospital.icu) of the step's summary row count the
// people PLACED there (records_writer.py:290-294), by the hospital's region.
__device__ __forceinline__ void jb_place_patient(const PlaceArgs& a, uint32_t p) {
  const uint32_t med = (uint32_t)a.pmed[p];  // (hospital group << 8) | local subgroup
  const uint32_t hg = med >> 8, lsg = med & 0xFu;
  jb_dyn_scatter(a, hg, lsg, p);
  if (a.summary && lsg >= 1u && lsg <= 2u)
    atomicAdd(&a.summary[JB_GINFO_REGION(a.grp_info[hg]) * JB_N_SUMMARY + (lsg == 2u ? SUM_CUR_ICU : SUM_CUR_HOSP)], 1u);
}

// Travel.get_commute_subgroup (travel.py:125-131) -> City.get_commute_subgroup (city.py:83-102) ->
// Station.get_commute_subgroup (station.py:52-53,73-76): the transport group, or -1.
__device__ __noinline__ int32_t jb_commute_transport(const PlaceArgs& a, uint32_t p) {
  const CommuteDev& c = a.cmt;
  if (c.inject) return c.inject[p];
  const int32_t code = c.person[p];
  if (code < 0) return -1;
  const JbPhilox4 r = jb_philox4x32_10(a.gid_base + a.int2ext[p], a.sp->step, JB_STREAM_COMMUTE, 0u, (uint32_t)a.sp->seed,
                                       (uint32_t)(a.sp->seed >> 32));
  uint32_t station;
  if (code & 1) {
    station = (uint32_t)code >> 1;
  } else {
    const uint32_t o = c.cs_off[code >> 1], n = c.cs_off[(code >> 1) + 1] - o;
    if (!n) return -1;
    const uint32_t j = (uint32_t)(jb_u52(r.x[0], r.x[1]) * (double)n);
    station = c.cs[o + (j < n ? j : n - 1)];
  }
  const uint32_t o = c.st_off[station], n = c.st_off[station + 1] - o;
  if (!n) return -1;
  const uint32_t j = (uint32_t)(jb_u52(r.x[2], r.x[3]) * (double)n);
  return (int32_t)c.st[o + (j < n ? j : n - 1)];
}

// Household._get_leisure_subgroup_for_person (household.py:110-123): the subgroup a visitor joins.
__device__ __forceinline__ uint32_t jb_visitor_subgroup(uint32_t age) { return age < 18u ? 0u : age <= 25u ? 1u : age < 65u ? 2u : 3u; }

// Leisure.get_subgroup_for_person_and_housemates (leisure.py:230-275) for social venues: does the
// person go out, which activity, which of its area's candidate venues
// (social_venue_distributor.py:257-266).  Returns the venue group (and its subgroup) or -1.
__device__ __noinline__ int32_t jb_leisure_venue(const PlaceArgs& a, uint32_t p, uint32_t& lsg) {
  const LeisureDev& z = a.lz;
  if (a.ch[p]) return -1;  // care-home residents stay (leisure.py:246-248)
  const VisitsDev& vz = a.vz;
  if (z.inject) {
    const int32_t g = z.inject[p];
    if (g < 0) return -1;
    if (vz.activity >= 0) {  // a residence visit: household (subgroup by age, household.py:110-123) or care home
      if ((uint32_t)g - vz.hh_first < vz.hh_count) { lsg = jb_visitor_subgroup(a.age[p]); return g; }
      if ((uint32_t)g - vz.ch_first < vz.ch_count) { lsg = vz.ch_lsg; return g; }
    }
    for (uint32_t k = 0; k < z.K; ++k)
      if ((uint32_t)g - z.sg_first[k] < z.sg_count[k]) { lsg = (uint32_t)z.lsg[k]; return g; }
    return -1;
  }
  const uint32_t K = z.K, tab = a.sp->leisure_table;
  if (!tab || tab > z.n_tables) return -1;
  const uint32_t age = min((uint32_t)a.age[p], 99u);
  const size_t row = ((size_t)(tab - 1) * a.R + a.pregion[p]) * 200 + (size_t)a.sex[p] * 100 + age;
  const uint32_t gid = a.gid_base + a.int2ext[p];
  const JbPhilox4 r0 = jb_philox4x32_10(gid, a.sp->step, JB_STREAM_LEISURE, 0u, (uint32_t)a.sp->seed, (uint32_t)(a.sp->seed >> 32));
  if (!(jb_u52(r0.x[0], r0.x[1]) < z.does[row])) return -1;
  const double u_act = jb_u52(r0.x[2], r0.x[3]);
  const double* cum = z.cum + row * K;
  uint32_t k = 0;
  while (k + 1 < K && !(u_act < cum[k])) ++k;
  if ((int32_t)k == vz.activity) {
    // ResidenceVisitsDistributor.get_leisure_group (residence_visits.py:297-331): the kinds of residence the person's
    // household is linked to, one of them by residence_type_probabilities, one of its candidates uniformly
    const uint32_t h = (uint32_t)vz.pres_group[p] - vz.hh_first;
    if (h >= vz.hh_count) return -1;
    const uint32_t n_hh = vz.hh_off[h + 1] - vz.hh_off[h];
    const uint32_t n_ch = vz.ch_off ? vz.ch_off[h + 1] - vz.ch_off[h] : 0u;
    if (!n_hh && !n_ch) return -1;
    const JbPhilox4 rv = jb_philox4x32_10(gid, a.sp->step, JB_STREAM_LEISURE, 1u, (uint32_t)a.sp->seed, (uint32_t)(a.sp->seed >> 32));
    const bool to_ch = !n_hh || (n_ch && !(jb_u52(rv.x[2], rv.x[3]) < vz.p_household));
    const uint32_t nc = to_ch ? n_ch : n_hh;
    uint32_t j = (uint32_t)(jb_u52(rv.x[0], rv.x[1]) * (double)nc);
    if (j >= nc) j = nc - 1;
    if (to_ch) { lsg = vz.ch_lsg; return (int32_t)vz.ch[vz.ch_off[h] + j]; }
    lsg = jb_visitor_subgroup(a.age[p]);
    return (int32_t)vz.hh[vz.hh_off[h] + j];
  }
  const uint32_t o = z.off[(size_t)z.parea[p] * K + k], n = z.off[(size_t)z.parea[p] * K + k + 1] - o;
  if (n == 0) return -1;  // get_leisure_group returns None: the person falls through to residence
  const JbPhilox4 r1 = jb_philox4x32_10(gid, a.sp->step, JB_STREAM_LEISURE, 1u, (uint32_t)a.sp->seed, (uint32_t)(a.sp->seed >> 32));
  uint32_t j = (uint32_t)(jb_u52(r1.x[0], r1.x[1]) * (double)n);
  if (j >= n) j = n - 1;
  lsg = (uint32_t)z.lsg[k];
  return (int32_t)z.venue[o + j];
}

// activity chosen for a person byte: pure function of (dead, severe, has-medical, fixed
// activities) and of the step's activity mask / policy flags -> 512-entry table per block
__device__ __forceinline__ uint32_t jb_place_key(uint32_t st, uint32_t has) {
  return (((st >> 3) & 0xFu) << 5) | (has & 0x1Fu);  // bits: INF DEAD SEVERE MED | has[5]
}

// FULL = false: neither leisure nor individual policies in this step -- the activity is a pure
// table lookup per byte (the common weekday-night / no-lockdown step keeps its 48 registers).
template <bool FULL>
__global__ void __launch_bounds__(256, FULL ? 1 : 6) k_place(const __grid_constant__ PlaceArgs a) {
  jb_stamp(1);
  __shared__ uint8_t s_act[512];      // 0..4 activity, 7 none (dead), 15 alive without any activity
  __shared__ uint32_t s_cnt[8];
  const uint32_t mask = a.sp->activity_mask, flags = a.sp->flags;
  const bool leisure_on = FULL && a.lz.K != 0 && (a.lz.inject != nullptr || a.sp->leisure_table != 0);
  const bool commute_on = FULL && (a.cmt.person != nullptr || a.cmt.inject != nullptr);
  for (uint32_t k = threadIdx.x; k < 512; k += blockDim.x) {
    const uint32_t bits = k >> 5, has0 = k & 0x1Fu;
    const bool dead = bits & 2u, severe = bits & 4u, med = bits & 8u;
    uint32_t act = JB_ACT_NONE;
    if (!dead) {
      uint32_t allowed = mask;
      if ((flags & JB_STEP_SEVERE_STAY_HOME) && severe) allowed &= (1u << JB_ACT_MEDICAL) | (1u << JB_ACT_RESIDENCE);
      const uint32_t cand = allowed & (has0 | (med ? 1u << JB_ACT_MEDICAL : 0u));
      act = cand ? (uint32_t)__ffs(cand) - 1u : 15u;  // hierarchy order == bit order
      // commute has no fixed subgroup either: drawn per commuter unless a medical facility comes first
      if (commute_on && (allowed & (1u << JB_ACT_COMMUTE)) && !(cand & (1u << JB_ACT_MEDICAL))) act |= 0x40u;
      // leisure has no fixed subgroup: it is drawn per person when nothing above it in the
      // hierarchy applies (activity_manager.py:380-398) -> flag bit 7, resolved below
      if (leisure_on && (allowed & (1u << JB_ACT_LEISURE)) && !(cand & ((1u << JB_ACT_LEISURE) - 1u))) act |= 0x80u;
    }
    s_act[k] = (uint8_t)act;
  }
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  // snapshot of the infection-slot high-water mark: slots below it are "old" for this step
  if (blockIdx.x == 0 && threadIdx.x == 0) a.counters[CNT_SLOTS_PREV] = a.counters[CNT_SLOTS_HW];
  __syncthreads();
  // eight 8-bit counters (one per activity code) packed in a 64-bit register
  unsigned long long cnt = 0ull;
  uint32_t noact = 0;
  const uint32_t n_chunks = (a.N + 15) / 16;
  auto one = [&](uint32_t p, uint32_t st, uint32_t has) -> uint32_t {
    uint32_t act;
    if (FULL && a.n_pol && !(st & JB_PS_DEAD)) {
      // individual policies decide per person (ages, tags, uniforms): the table's logic, inline
      uint32_t allowed = mask;
      if ((flags & JB_STEP_SEVERE_STAY_HOME) && (st & JB_PS_SEVERE)) allowed &= (1u << JB_ACT_MEDICAL) | (1u << JB_ACT_RESIDENCE);
      else allowed = jb_apply_policies(a, p, st, allowed);
      const uint32_t cand = allowed & ((has & 0x1Fu) | ((st & JB_PS_MEDICAL) ? 1u << JB_ACT_MEDICAL : 0u));
      act = cand ? (uint32_t)__ffs(cand) - 1u : 15u;
      if (leisure_on && (allowed & (1u << JB_ACT_LEISURE)) && !(cand & ((1u << JB_ACT_LEISURE) - 1u))) act |= 0x80u;
      if (commute_on && (allowed & (1u << JB_ACT_COMMUTE)) && !(cand & (1u << JB_ACT_MEDICAL))) act |= 0x40u;
    } else {
      act = s_act[jb_place_key(st, has)];
    }
    if (FULL && (act & 0x40u)) {  // commute draw (above primary activity, leisure and residence)
      const int32_t g = jb_commute_transport(a, p);
      if (g >= 0) { act = JB_ACT_COMMUTE; jb_dyn_scatter(a, (uint32_t)g, 0u, p); }
      else act &= ~0x40u;
    }
    if (FULL && (act & 0x80u)) {  // leisure draw
      act &= 0x7Fu;
      uint32_t lsg = 0;
      const int32_t g = jb_leisure_venue(a, p, lsg);
      if (g >= 0) {
        act = JB_ACT_LEISURE;
        if (!a.vis_defer) {
          jb_dyn_scatter(a, (uint32_t)g, lsg, p);
        } else {
          // residence visits: whether the person really goes depends on who visits whom -- remember the destination,
          // and tell a household it is being visited (by the caller's index of the visitor: world order)
          a.vz.choice[p] = (lsg << 28) | (uint32_t)g;
          a.vz.list[atomicAdd(&a.counters[CNT_VIS_LIST], 1u)] = p;
          if ((uint32_t)g - a.vz.hh_first < a.vz.hh_count) atomicMin(&a.vz.hv[0][(uint32_t)g - a.vz.hh_first], a.int2ext[p]);
        }
      }
    }
    if (act == 15u) { ++noact; act = JB_ACT_NONE; }
    cnt += 1ull << (8 * act);
    if (act == JB_ACT_MEDICAL) jb_place_patient(a, p);  // rare: the patient goes into its ward's bin
    const uint32_t var = a.pvar ? (uint32_t)a.pvar[p] : 0u;
    if (FULL && a.mstep) {  // this step's byte of every mirrored person (policies / leisure / commute decide per person)
      const uint32_t slot = a.pmslot[p];
      if (slot != JB_MB_NONE) a.mstep[slot] = (uint8_t)(jb_mirror_byte(st, var) | (act != JB_ACT_PRIMARY ? JB_MB_ABSENT : 0u));
    }
    if (!FULL && a.pmir) {  // persistent mirror: push the byte if it changed since it was last pushed
      const uint32_t mb = jb_mirror_byte(st, var);
      if (a.pmir[p] != mb) {
        a.pmir[p] = (uint8_t)mb;
        const uint32_t slot = a.pmslot[p];
        if (slot != JB_MB_NONE) a.mstate[slot] = (uint8_t)mb;
      }
    }
    // the step's hot byte: activity | INFECTED | susceptibility code | variant
    return act | (st & JB_PS_INFECTED) | ((st & JB_PST_CODE_MASK) << 4) | (var << 6);
  };
  uint32_t iters = 0;
  // SWAR path: byte-wide per-activity counters of the words handled since the last flush, chunks handled, persons without a candidate
  uint32_t pl0 = 0, pl1 = 0, pl2 = 0, pl3 = 0, n_none = 0, n_swar = 0;
  auto flush_planes = [&]() {
    const uint32_t c0 = (pl0 * 0x01010101u) >> 24, c1 = (pl1 * 0x01010101u) >> 24, c2 = (pl2 * 0x01010101u) >> 24, c3 = (pl3 * 0x01010101u) >> 24;
    const uint32_t c4 = n_swar * 16u - n_none - c0 - c1 - c2 - c3;
    cnt += (unsigned long long)c0 | ((unsigned long long)c1 << 8) | ((unsigned long long)c2 << 16) | ((unsigned long long)c3 << 24) |
           ((unsigned long long)c4 << 32);
    pl0 = pl1 = pl2 = pl3 = 0; n_none = 0; n_swar = 0;
  };
  if (FULL) {
    // one person per thread: policies, commute and leisure walk per-person tables and draw Philox
    // numbers -- a dependent chain per person that 16 persons per thread would serialise
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < a.N; p += gridDim.x * blockDim.x) {
      a.phot[p] = (uint8_t)one(p, a.pstate[p], a.phas[p]);
      if (++iters == 240) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(&s_cnt[k], (uint32_t)((cnt >> (8 * k)) & 0xFFull));
        cnt = 0ull; iters = 0;
      }
    }
  } else
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += gridDim.x * blockDim.x) {
    const uint32_t p0 = c * 16;
    if (p0 + 16 <= a.N) {
      const uint4 sv = *reinterpret_cast<const uint4*>(a.pstate + p0);
      const uint4 hv = *reinterpret_cast<const uint4*>(a.phas + p0);
      {  // this thread's next chunk starts its trip from DRAM now (the even lane asks for the pair's 32-byte sector)
        const size_t pn = (size_t)p0 + (size_t)gridDim.x * blockDim.x * 16u;
        if (!(threadIdx.x & 1) && pn < a.N) { jb_prefetch_l2(a.pstate + pn); jb_prefetch_l2(a.phas + pn); }
      }
      uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
      const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
      if (!a.pvar) {
        ++n_swar;
        // Four persons per 32-bit word, no table: candidate activities = (fixed subgroups | medical)
        // & the step's mask, restricted for severe stay-at-home cases, none for the dead; the chosen
        // activity is the lowest set bit of each byte (hierarchy order == bit order).
        const uint32_t M4 = (mask & 0x1Fu) * 0x01010101u;
        const uint32_t keep_sev = ((1u << JB_ACT_MEDICAL) | (1u << JB_ACT_RESIDENCE)) * 0x01010101u;
        if (a.pmir) {
          // persistent mirror of the primary-activity supergroups: state bytes that changed since they were last
          // pushed (new infections, recoveries, hospital, death, vaccination: a few per cent of the people
          // between two primary-activity steps) go to their slot
          const uint4 pm = *reinterpret_cast<const uint4*>(a.pmir + p0);
          const uint32_t pw[4] = {pm.x, pm.y, pm.z, pm.w};
          uint32_t nw[4];
          bool any = false;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            nw[q] = sw[q] & 0x7B7B7B7Bu;
            uint32_t diff = nw[q] ^ pw[q];
            while (diff) {
              const uint32_t b = (uint32_t)(__ffs(diff) - 1) >> 3;
              diff &= ~(0xFFu << (8 * b));
              any = true;
              const uint32_t slot = a.pmslot[p0 + q * 4 + b];
              if (slot != JB_MB_NONE) a.mstate[slot] = (uint8_t)(nw[q] >> (8 * b));
            }
          }
          if (any) *reinterpret_cast<uint4*>(a.pmir + p0) = make_uint4(nw[0], nw[1], nw[2], nw[3]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t st4 = sw[q];
          const uint32_t med4 = (st4 >> 6) & 0x01010101u, dead4 = (st4 >> 4) & 0x01010101u, sev4 = (st4 >> 5) & 0x01010101u;
          uint32_t cand4 = ((hw[q] & 0x1F1F1F1Fu) | med4) & M4;
          if (flags & JB_STEP_SEVERE_STAY_HOME) cand4 &= ~((sev4 * 0xFFu) & ~keep_sev);
          cand4 &= ~(dead4 * 0xFFu);
          const uint32_t low4 = cand4 & ~((cand4 | 0x80808080u) - 0x01010101u);  // lowest set bit of every byte
          const uint32_t none4 = (~(((low4 & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | low4) >> 7) & 0x01010101u;  // 1 where no candidate
          uint32_t act4 = (((low4 >> 1) | (low4 >> 3)) & 0x01010101u) | ((((low4 >> 2) | (low4 >> 3)) & 0x01010101u) << 1) |
                          (((low4 >> 4) & 0x01010101u) << 2);
          act4 |= none4 * 7u;  // JB_ACT_NONE
          noact += __popc(none4 & ~dead4);
          // persons per activity: four byte-wide counters per activity (one per person of the word); the residence,
          // the last and most common activity, is what remains of the chunk's 16 persons
          pl0 += low4 & 0x01010101u; pl1 += (low4 >> 1) & 0x01010101u; pl2 += (low4 >> 2) & 0x01010101u; pl3 += (low4 >> 3) & 0x01010101u;
          n_none += __popc(none4);
          uint32_t medm = low4 & 0x01010101u;  // rare: patients go into their ward's bin
          while (medm) {
            const uint32_t b = (uint32_t)(__ffs(medm) - 1) >> 3;
            medm &= medm - 1;
            jb_place_patient(a, p0 + q * 4 + b);
          }
          sw[q] = act4 | (st4 & 0x08080808u) | ((st4 & 0x03030303u) << 4);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t o = 0;
#pragma unroll
          for (int b = 0; b < 4; ++b)
            o |= one(p0 + q * 4 + b, (sw[q] >> (8 * b)) & 0xFFu, (hw[q] >> (8 * b)) & 0xFFu) << (8 * b);
          sw[q] = o;
        }
      }
      *reinterpret_cast<uint4*>(a.phot + p0) = make_uint4(sw[0], sw[1], sw[2], sw[3]);
    } else {
      for (uint32_t p = p0; p < a.N; ++p) a.phot[p] = (uint8_t)one(p, a.pstate[p], a.phas[p]);
    }
    if (++iters == 15) {  // 15 * 16 = 240 < 256: flush the 8-bit counters
      flush_planes();
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&s_cnt[k], (uint32_t)((cnt >> (8 * k)) & 0xFFull));
      cnt = 0ull; iters = 0;
    }
  }
  if (!FULL) flush_planes();
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

// ---- residence visits: who really goes (household.py:124-149,202-205; activity_manager.py:275-292,380-383) ------------
// The reference walks the people in world order.  A person whose household has ALREADY received a visitor when its turn
// comes stays home without drawing a leisure activity ("not executed"); a visitor who chose household t marks t as being
// visited (its residents who are out doing leisure come back, the ones still to come stay).  With one Philox stream per
// person every draw is a function of (person, step), so the only order-dependent fact is which draws are executed:
//     executed(i) = no executed visitor k < i chose i's household.
// k_place has recorded every destination and, per household, the smallest index of ANY visitor (hv[0]); each pass below
// recomputes the minima over the visitors that are executed according to the previous pass.  A person's status is final
// after as many passes as its chain of "visited by somebody who is visited by ..." is long.
__global__ void __launch_bounds__(256) k_visit_iter(const __grid_constant__ VisitsDev z, const uint32_t* __restrict__ counters,
                                                    const uint32_t* __restrict__ int2ext, int src, int dst, int clr) {
  const uint32_t n = counters[CNT_VIS_LIST];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t p = z.list[i];
    const uint32_t t = (z.choice[p] & 0x07FFFFFFu) - z.hh_first;
    if (t >= z.hh_count) continue;  // a venue or a care home: nobody is kept home by it
    z.hv[clr][t] = JB_VIS_NONE;     // the buffer the NEXT pass accumulates into
    const uint32_t ext = int2ext[p];
    const uint32_t h = (uint32_t)z.pres_group[p] - z.hh_first;
    if (h >= z.hh_count || !(z.hv[src][h] < ext)) atomicMin(&z.hv[dst][t], ext);
  }
}

// Everybody with a destination is placed: home if its household is being visited (whether it was kept from drawing or
// fetched back), else into the bin of its venue / care home, or -- visitors of a household -- left for k_visit_scatter.
// The first executed visitor of a household opens its entry in the visited list.
__global__ void __launch_bounds__(256) k_visit_final(const __grid_constant__ PlaceArgs a, int fin) {
  const VisitsDev& z = a.vz;
  const uint32_t n = a.counters[CNT_VIS_LIST];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t p = z.list[i];
    const uint32_t c = z.choice[p], g = c & 0x07FFFFFFu, lsg = c >> 28;
    const uint32_t ext = a.int2ext[p];
    const uint32_t h = (uint32_t)z.pres_group[p] - z.hh_first, t = g - z.hh_first;
    const uint32_t first_home = h < z.hh_count ? z.hv[fin][h] : JB_VIS_NONE;
    const bool executed = !(first_home < ext);
    if (t < z.hh_count && executed && z.hv[fin][t] == ext) {  // Household.being_visited = True
      const uint32_t v = atomicAdd(&a.counters[CNT_VIS_N], 1u);
      if (v < z.vis_cap) { z.vis_group[v] = g; z.vslot[t] = v; }
      else atomicAdd(&a.counters[CNT_OVERFLOW], 1u);
    }
    if (first_home == JB_VIS_NONE) {  // goes
      if (t >= z.hh_count) { jb_dyn_scatter(a, g, lsg, p); z.choice[p] = c | JB_VIS_DONE; }
    } else {                          // stays / comes back home
      a.phot[p] = (uint8_t)((a.phot[p] & ~JB_PS_ACT_MASK) | JB_ACT_RESIDENCE);
      atomicSub(&a.counters[CNT_ACT0 + JB_ACT_LEISURE], 1u);
      atomicAdd(&a.counters[CNT_ACT0 + JB_ACT_RESIDENCE], 1u);
      z.choice[p] = c | JB_VIS_DONE;
    }
  }
}

// Visitors of households join the visited household's entry: (subgroup << 28 | caller's index), sorted by the group kernel.
__global__ void __launch_bounds__(256) k_visit_scatter(const __grid_constant__ VisitsDev z, uint32_t* __restrict__ counters,
                                                       const uint32_t* __restrict__ int2ext) {
  const uint32_t n = counters[CNT_VIS_LIST];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t p = z.list[i], c = z.choice[p];
    if (c & JB_VIS_DONE) continue;  // placed already (home, venue or care home)
    const uint32_t v = z.vslot[(c & 0x07FFFFFFu) - z.hh_first];
    if (v == JB_VIS_NONE) continue;  // (visited list full: already reported)
    const uint32_t k = atomicAdd(&z.vis_cnt[v], 1u);
    if (k < JB_VIS_BIN) z.vis_bin[(size_t)v * JB_VIS_BIN + k] = (c & 0xF0000000u) | int2ext[p];
    else atomicAdd(&counters[CNT_OVERFLOW], 1u);
  }
}

// End of the step: everything the step's visitors touched goes back to "nobody" (no full-array clears per step).
__global__ void __launch_bounds__(256) k_visit_cleanup(const __grid_constant__ VisitsDev z, const uint32_t* __restrict__ counters) {
  const uint32_t n = counters[CNT_VIS_LIST], nv = min(counters[CNT_VIS_N], z.vis_cap);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < max(n, nv); i += gridDim.x * blockDim.x) {
    if (i < nv) {
      z.vis_cnt[i] = 0u;
      z.vslot[z.vis_group[i] - z.hh_first] = JB_VIS_NONE;
    }
    if (i < n) {
      const uint32_t p = z.list[i];
      const uint32_t t = (z.choice[p] & 0x07FFFFFFu) - z.hh_first;
      if (t < z.hh_count) z.hv[0][t] = z.hv[1][t] = z.hv[2][t] = JB_VIS_NONE;
      z.choice[p] = JB_VIS_NONE;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Stage B: interaction.
// ------------------------------------------------------------------------------------------
// Global person id (the Philox counter word): residents are renumbered internally
// (int2ext maps back to the caller's index), halo persons carry their home domain's id.
__device__ __forceinline__ uint32_t jb_gid(const GroupArgs& a, uint32_t p) {
  return p < a.N ? a.gid_base + a.int2ext[p] : a.halo_gid[p - a.N];
}

// Append one new infection (and, when events are recorded, where and by whom).
__device__ __forceinline__ void jb_emit(const GroupArgs& a, uint32_t p, uint32_t v, uint32_t g, uint32_t infector) {
  uint32_t i = atomicAdd(&a.counters[CNT_NEWINF], 1u);
  if (i < a.newinf_cap) {
    a.newinf_member[i] = p;
    a.newinf_var[i] = (uint8_t)v;
    if (a.ev_group) { a.ev_infected[i] = p; a.ev_var[i] = (uint8_t)v; a.ev_group[i] = g; a.ev_infector[i] = infector; }
  } else {
    atomicAdd(&a.counters[CNT_OVERFLOW], 1u);
  }
}

// Bernoulli draw + variant choice for one susceptible (interaction.py:610-641).  Returns the
// variant the person caught, or -1.
// `gid` = the person's global id (jb_gid, or the slot-ordered copy of it in a mirrored supergroup).
template <int VM>
__device__ __forceinline__ int jb_draw_core(const GroupArgs& a, uint32_t gid, const double* lam_v, double lam,
                                            uint32_t ordinal) {
  double u;
  if (a.inject) {
    u = ordinal < a.sp->n_inject ? a.inject[ordinal] : 2.0;
  } else {
    u = jb_bernoulli_uniform(a.sp->seed, a.sp->step, gid);
  }
  // 1 - exp(-lam) <= lam: a uniform above lam (plus a margin for the rounding of exp and of the subtraction, < 4e-16)
  // cannot succeed -- the common case, decided without the exponential; the outcome is the same either way
  if (!(u < lam + 1e-15)) return -1;
  if (!(u < 1.0 - exp(-lam))) return -1;
  uint32_t v = 0;
  if (VM > 1 && a.V > 1) {
    // np.random.choice(ids, p=lam_v/lam): inverse CDF on a second, independent uniform
    double u2 = jb_variant_uniform(a.sp->seed, a.sp->step, gid) * lam;
    double c = 0.0;
    v = a.V - 1;
    for (int k = 0; k < a.V; ++k) {
      c += lam_v[k];
      if (u2 < c) { v = k; break; }
    }
  }
  return (int)v;
}

// The two uniforms of an infection's blame (_blame_subgroup / _blame_individuals,
// interaction.py:643-662): np.random.choice(p=...) picks the first index whose cumulative
// probability exceeds the uniform.
__device__ __forceinline__ void jb_blame_uniforms(const GroupArgs& a, uint32_t p, double& u_sub, double& u_ind) {
  const JbPhilox4 r = jb_philox4x32_10(jb_gid(a, p), a.sp->step, JB_STREAM_BLAME, 0u, (uint32_t)a.sp->seed,
                                       (uint32_t)(a.sp->seed >> 32));
  u_sub = jb_u52(r.x[0], r.x[1]);
  u_ind = jb_u52(r.x[2], r.x[3]);
}

#include "jb_groups.cuh"
#include "jb_span.cuh"

// Fix-up pass of the table-driven placement for policies whose subjects are a static, short list of
// people (LimitLongCommute: the long-distance commuters, individual_policies.py:773-824): everyone is
// placed by the fast path, then the listed people are re-evaluated with the full policy logic.  Nobody
// else can be touched by such a policy and nobody else draws, so the result equals the full path's.
__global__ void __launch_bounds__(256) k_place_subjects(const __grid_constant__ PlaceArgs a, const uint32_t* __restrict__ subjects,
                                                        uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t p = subjects[i];
  const uint32_t st = a.pstate[p];
  if (st & JB_PS_DEAD) return;
  const uint32_t mask = a.sp->activity_mask, flags = a.sp->flags, has = a.phas[p];
  uint32_t allowed = mask;
  if ((flags & JB_STEP_SEVERE_STAY_HOME) && (st & JB_PS_SEVERE)) allowed &= (1u << JB_ACT_MEDICAL) | (1u << JB_ACT_RESIDENCE);
  else allowed = jb_apply_policies(a, p, st, allowed);
  const uint32_t cand = allowed & ((has & 0x1Fu) | ((st & JB_PS_MEDICAL) ? 1u << JB_ACT_MEDICAL : 0u));
  const uint32_t hot = a.phot[p], old_act = hot & JB_PS_ACT_MASK;
  uint32_t act = cand ? (uint32_t)__ffs(cand) - 1u : (uint32_t)JB_ACT_NONE;
  if (a.mstate && (mask & (1u << JB_ACT_PRIMARY)) && act != JB_ACT_PRIMARY) {
    // absent from its mirrored slot in this step only: the invalidated `pmir` makes the next table-driven placement
    // push the person's byte again
    const uint32_t slot = a.pmslot[p];
    if (slot != JB_MB_NONE) { a.mstate[slot] = (uint8_t)(a.mstate[slot] | JB_MB_ABSENT); a.pmir[p] = 0xFFu; }
  }
  if (act == old_act) return;
  // (a medical facility precedes every activity a policy can remove: ward bins are unaffected)
  if (old_act < JB_N_ACT) atomicSub(&a.counters[CNT_ACT0 + old_act], 1u);
  if (act < JB_N_ACT) atomicAdd(&a.counters[CNT_ACT0 + act], 1u);
  else atomicAdd(&a.counters[CNT_NOACT], 1u);
  a.phot[p] = (uint8_t)((hot & ~JB_PS_ACT_MASK) | act);
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
  uint8_t* pstate; int8_t* ptag; double* trans; uint32_t NP; int32_t* pmed; int32_t* pslot; uint8_t* phas;  // trans: [2][NP], written half = tpar ^ 1
  uint32_t* counters;
  uint32_t* summary;        // [R][JB_N_SUMMARY]
  const uint32_t* grp_info; // region of a hospital
  int R;
  int old_only;             // 1: only the slots that existed at step start (k_newinf updates its own)
  // health events (null unless JB_STEP_RECORD_EVENTS)
  uint32_t* hev_person; uint32_t* hev_kind; int32_t* hev_group; uint32_t hev_cap;
  const int32_t* pres_group;
};

__device__ __forceinline__ void jb_health_event(const HealthArgs& a, uint32_t p, uint32_t kind, int32_t group) {
  const uint32_t i = atomicAdd(&a.counters[CNT_HEV], 1u);
  if (i < a.hev_cap) { a.hev_person[i] = p; a.hev_kind[i] = kind; a.hev_group[i] = group; }
  else atomicAdd(&a.counters[CNT_OVERFLOW], 1u);
}

struct HealthAcc {
  uint32_t inf, hosp, icu;  // currently infected / in hospital / in intensive care
};

#define JB_SUM_SMEM_REGIONS 64

// Gamma transmission profile (transmission.py:47-75): probability(t) = norm * gamma.pdf(t, a, loc, scale)
//   = norm / (Gamma(a) scale) * z^(a-1) e^(-z),  z = (t - loc) / scale.
// The factor in front does not change during an infection: it is computed once, when the infection is created
// (row 4 of inf_par, which the gamma profile does not otherwise use), and the health update of every step is left
// with one logarithm and one exponential per infection instead of tgamma + pow + exp (the update is bound by fp64
// arithmetic: ~3.5 M infectious people at 56 M).
__device__ __forceinline__ double jb_gamma_front(double norm, double a, double scale) { return norm * (1.0 / tgamma(a)) / scale; }
__device__ __forceinline__ double jb_gamma_profile(double front, double x, double a, double loc, double scale) {
  if (x < loc) return 0.0;
  const double z = (x - loc) / scale;
  if (!(z > 0.0)) return front * pow(z, a - 1.0);  // z == 0: 0, 1 or inf depending on a -- pow's business
  return front * exp((a - 1.0) * log(z) - z);
}

// One thread per infection slot.  Everything the update needs about the PERSON (region, closest
// hospital, ward, severe flag) is mirrored in the slot, so the only per-step access indexed by
// person is the write of the new transmission probability; the person's state byte, ward and
// tag are touched only when they change.
// Update of one infection slot.  Everything the update needs about the PERSON (region, closest
// hospital, ward, severe flag) is mirrored in the slot, so the only per-step access indexed by
// person is the write of the new transmission probability; the person's state byte, ward and
// tag are touched only when they change.  `sum_base` = per-region counters (shared or global).
__device__ __forceinline__ void jb_health_slot(const HealthArgs& a, double* trans, uint32_t s, double t_now, uint32_t flags,
                                               uint32_t* sum_base, HealthAcc& acc) {
  const DiseaseMasks D = a.masks;
  // Everything the slot holds is requested up front, in ONE round trip to DRAM: left where they are used, the loads
  // sit behind the branch on the person and behind the store of the transmission probability (which the compiler
  // must assume to alias them) and reach DRAM one after the other -- five latencies per slot instead of one.
  const int32_t p = a.inf_person[s];
  const double start_s = a.inf_start[s], next_s = a.inf_next[s];
  const double par0 = a.inf_par[s], par1 = a.inf_par[(size_t)1 * a.C + s], par2 = a.inf_par[(size_t)2 * a.C + s];
  const double par3 = D.trans_type == JB_TRANS_XNEXP ? a.inf_par[(size_t)3 * a.C + s] : 0.0;
  const double par4 = D.trans_type != JB_TRANS_CONSTANT ? a.inf_par[(size_t)4 * a.C + s] : 0.0;
  int stage = a.inf_stage[s];
  int tag = a.inf_tag[s];
  const int nst = a.inf_nst[s];
  const uint32_t region_s = a.inf_region[s];
  const int32_t med_old = a.inf_med[s];
  const uint32_t sev_old = a.inf_sev[s];
  if (p < 0) return;
  const double t = t_now - start_s;
  // An infection that moves on to its next stage in this step (a few per cent of them, but somebody in almost every
  // warp) needs the stage's tag, the time of the stage after it and the person's state byte: a second trip to DRAM,
  // requested here so that it overlaps the profile's arithmetic instead of following the store below.
  const bool advance = stage + 1 < nst && t > next_s;
  int tag_adv = tag;
  double next_adv = 1e300;
  uint32_t st_adv = 0;
  if (advance) {
    tag_adv = a.traj_tag[(size_t)(stage + 1) * a.C + s];
    if (stage + 2 < JB_MAX_STAGES) next_adv = a.traj_time[(size_t)(stage + 2) * a.C + s];
    st_adv = a.pstate[p];
  }
  // transmission.update_infection_probability (infection.py:91)
  if (D.trans_type == JB_TRANS_GAMMA) {
    trans[p] = jb_gamma_profile(par4, t, par0, par1, par2);
  } else if (D.trans_type == JB_TRANS_XNEXP) {
    const double tfi = par2;
    double prob = 0.0;
    if (t > tfi) {
      const double tau = (t - tfi) / par4;
      prob = par3 * (pow(tau, par0) * exp(-tau / par1));
    }
    trans[p] = prob;
  } else {
    trans[p] = par0;  // constant profile: carried over into the buffer the next step reads
  }
  // symptoms.update_trajectory_stage: at most one stage per call (symptoms.py:152-154)
  bool changed = false;
  if (advance) {
    ++stage;
    const int tag_old = tag;
    tag = tag_adv;
    if (a.hev_person && tag != tag_old) jb_health_event(a, (uint32_t)p, JB_HEV_SYMPTOMS, tag);
    a.inf_stage[s] = (uint8_t)stage;
    a.inf_tag[s] = (int8_t)tag;
    a.inf_next[s] = next_adv;
    changed = true;
  }
  const int bit = JB_TAG_BIT(tag);
  uint32_t* sum = sum_base + (size_t)region_s * JB_N_SUMMARY;
  // what happens in a hospital is counted for the HOSPITAL's region (records_writer.py:272-294,369-373)
  auto hsum = [&](int32_t hg) -> uint32_t* { return sum_base + (size_t)JB_GINFO_REGION(a.grp_info[hg]) * JB_N_SUMMARY; };
  // Hospitalisation.apply (medical_care_policies.py:413-526; hospital.py:62-122)
  const bool was_med = med_old >= 0;
  int32_t med_new = med_old;
  if (flags & JB_STEP_HOSPITALISATION) {
    if (bit & D.hosp_mask) {
      const int32_t hg = was_med ? (med_old >> 8) : a.inf_hosp[s];
      if (hg >= 0) med_new = (hg << 8) | ((bit & D.icu_mask) ? 2 : 1);
      if (hg >= 0) {
        // "ward_admitted" -> hospital_admissions, "icu_transferred" -> icu_admissions; a direct ICU
        // admission is not recorded by the reference (medical_care_policies.py:463-475)
        if (!was_med && (med_new & 0xFF) == 1) {
          atomicAdd(&hsum(hg)[SUM_ADMISSIONS], 1u);
          if (a.hev_person) jb_health_event(a, (uint32_t)p, JB_HEV_HOSPITAL_ADMISSION, hg);
        } else if (was_med && (med_old & 0xFF) == 1 && (med_new & 0xFF) == 2) {
          atomicAdd(&hsum(hg)[SUM_ICU_ADMISSIONS], 1u);
          if (a.hev_person) jb_health_event(a, (uint32_t)p, JB_HEV_ICU_ADMISSION, hg);
        }
      }
    } else if (was_med && !(bit & D.dead_mask)) {
      if (a.hev_person) jb_health_event(a, (uint32_t)p, JB_HEV_DISCHARGE, med_old >> 8);
      med_new = -1;
    }
  }
  const uint32_t sev_new = (bit & D.severe_mask) ? 1u : 0u;
  const bool rec = (bit & D.rec_mask) != 0, dead = !rec && (bit & D.dead_mask) != 0;
  if (rec || dead || changed || med_new != med_old || sev_new != sev_old) {
    uint32_t st = advance ? st_adv : (uint32_t)a.pstate[p];  // (nobody else writes this person's byte during the update)
    st = (med_new >= 0) ? (st | JB_PS_MEDICAL) : (st & ~JB_PS_MEDICAL);
    st = sev_new ? (st | JB_PS_SEVERE) : (st & ~JB_PS_SEVERE);
    if (med_new != med_old) { a.pmed[p] = med_new; a.inf_med[s] = med_new; }
    a.inf_sev[s] = (uint8_t)sev_new;
    if (rec) {  // Epidemiology.recover (epidemiology.py:197-215)
      st &= ~(JB_PS_INFECTED | JB_PS_SEVERE);
      a.inf_person[s] = -1;
      a.pslot[p] = -1;
      trans[p] = 0.0;
    } else if (dead) {  // bury_the_dead (epidemiology.py:168-195)
      st &= ~(JB_PS_INFECTED | JB_PS_SEVERE | JB_PS_MEDICAL);
      st |= JB_PS_DEAD;
      a.inf_person[s] = -1;
      a.pslot[p] = -1;
      a.pmed[p] = -1;
      a.phas[p] = 0;
      trans[p] = 0.0;
    }
    // ptag mirrors person.infection.tag; -1 ("healthy") once the infection object is gone
    a.ptag[p] = (rec || dead) ? (int8_t)-1 : (int8_t)tag;
    a.pstate[p] = (uint8_t)st;
  }
  if (a.hev_person && (rec || dead))
    jb_health_event(a, (uint32_t)p, rec ? JB_HEV_RECOVERY : JB_HEV_DEATH,
                    rec ? -1 : (med_new >= 0 ? (med_new >> 8) : a.pres_group[p]));
  if (rec) {
    atomicAdd(&a.counters[CNT_RECOVERED_STEP], 1u);
    atomicAdd(&sum[SUM_RECOVERED], 1u);
  } else if (dead) {
    atomicAdd(&a.counters[CNT_DEATHS_STEP], 1u);
    atomicAdd(&a.counters[CNT_DEAD_TOTAL], 1u);
    atomicAdd(&sum[SUM_DEATHS], 1u);
    if (med_new >= 0) atomicAdd(&hsum(med_new >> 8)[SUM_HOSP_DEATHS], 1u);  // died in its ward (epidemiology.py:176-179)
  } else {
    ++acc.inf;
    atomicAdd(&sum[SUM_CUR_INFECTED], 1u);
    if (med_new >= 0) {
      ++acc.hosp;
      if (bit & D.icu_mask) ++acc.icu;
    }
  }
}

// One thread per infection slot; `bid` of `nb` blocks (the caller may be a fused kernel).
__device__ __forceinline__ void jb_health_body(const HealthArgs& a, uint32_t bid, uint32_t nb) {
  // per-region counters are accumulated in shared memory and flushed once per block: a global
  // atomic per infected person on nine addresses serialises in L2
  __shared__ uint32_t s_sum[JB_SUM_SMEM_REGIONS * JB_N_SUMMARY];
  const bool sum_smem = a.R <= JB_SUM_SMEM_REGIONS;
  if (sum_smem) {
    for (int i = threadIdx.x; i < a.R * JB_N_SUMMARY; i += blockDim.x) s_sum[i] = 0;
    __syncthreads();
  }
  const uint32_t hw = min(a.old_only ? a.counters[CNT_SLOTS_PREV] : a.counters[CNT_SLOTS_HW], a.C);
  const double t_now = a.sp->t_now;
  const uint32_t flags = a.sp->flags;
  double* const trans = jb_trans_write(a.trans, a.NP, a.sp);
  HealthAcc acc = {0u, 0u, 0u};
  for (uint32_t s = bid * blockDim.x + threadIdx.x; s < hw; s += nb * blockDim.x)
    jb_health_slot(a, trans, s, t_now, flags, sum_smem ? s_sum : a.summary, acc);
  // block-level reduction of the "current" counters
  uint32_t c_inf = __reduce_add_sync(JB_FULL, acc.inf);
  uint32_t c_hosp = __reduce_add_sync(JB_FULL, acc.hosp);
  uint32_t c_icu = __reduce_add_sync(JB_FULL, acc.icu);
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

__global__ void __launch_bounds__(256, 4) k_health(const __grid_constant__ HealthArgs a) {
  jb_stamp(5);
  jb_health_body(a, blockIdx.x, gridDim.x);
}

// ------------------------------------------------------------------------------------------
// Stage C1: create infections (InfectionSelector._make_infection).
// ------------------------------------------------------------------------------------------
struct NewInfArgs {
  const uint32_t* member;   // person indices (local, or >= N for halo: skipped here)
  const uint8_t* variant;
  const uint32_t* count_dev;  // number of entries (device) or null -> count_host
  const uint32_t* first_dev;  // first entry of the list to walk (device; residents infected abroad follow the step's own list) or null = 0
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
  uint8_t* pstate; uint8_t* pvar; int8_t* ptag; double* susc; int32_t* pslot;
  double* trans;            // [2][NP]; half sp->tpar ^ 1 is the one the NEXT step reads, half sp->tpar the current one
                            // (written as well when no step is running: seeding)
  int32_t* inf_person; double* inf_start; double* inf_par; double* inf_next; double* traj_time;
  uint8_t* inf_stage; uint8_t* inf_nst; uint8_t* inf_var; int8_t* inf_tag; int8_t* inf_maxtag; int8_t* traj_tag;
  uint8_t* inf_region; int32_t* inf_hosp; int32_t* inf_med; uint8_t* inf_sev;  // slot-resident person facts
  const uint32_t* psa; const int32_t* sa_hospital; const int32_t* pmed;
  uint32_t* counters;
  uint32_t* summary;        // [R][JB_N_SUMMARY]
  const uint8_t* pregion;
  uint8_t* halo_ret;        // per halo person: variant+1 if infected here (or null)
  int mutate;               // 1: Mutation event -- the listed people are infected; only their transmission is replaced
  uint32_t mutate_variant;
  const double* effmult;    // [V][N] immunity.effective_multiplier_dict, or null (no vaccination)
  int fuse_health;          // 1 inside a step: the creating lane also runs the slot's first health update
  HealthArgs h;
};

// Sixteen sample slots per new infection, each an independent sample with its own Philox stream
// (JB_STREAM_SAMPLE0 + slot): 0 max_severity; 1..8 stage s = slot-1; 9..14 transmission parameter k = slot-9; 15 the
// second `alpha()` call of the xnexp profile.  The disease tables (6 KB, walked with data-dependent indices) travel as
// a kernel parameter: constant bank reads instead of a per-block staging copy behind a barrier.
// Assemble one infection from its samples (InfectionSelector._make_infection, infection_selector.py:146-332) and give
// it its first health update; run by one lane per infection.
__device__ __forceinline__ void jb_newinf_assemble(const NewInfArgs& a, const DiseaseDev& D, uint32_t p, uint32_t var, uint32_t st,
                                                   uint32_t slot, uint32_t region_p, int32_t hosp_p, int32_t med_p, int idx, int k,
                                                   int nst, const double* tstage, const double* tp, double* trans_next,
                                                   double* trans_cur) {
  if (slot >= a.C) { atomicAdd(&a.counters[CNT_OVERFLOW], 1u); return; }
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
    if (!a.mutate) {  // (a mutation keeps the person's symptoms)
      a.traj_time[(size_t)s * a.C + slot] = tt;
      a.traj_tag[(size_t)s * a.C + slot] = (int8_t)tg;
    }
  }
  // transmission (infection_selector.py:280-332)
  double par[JB_N_TPAR] = {0, 0, 0, 0, 0};
  if (D.trans_type == JB_TRANS_GAMMA) {
    par[0] = tp[1]; par[1] = tp[3] + t_exposed; par[2] = 1.0 / tp[2]; par[3] = tp[0];
    par[4] = jb_gamma_front(par[3], par[0], par[2]);
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
  if (a.mutate) {  // event/mutation.py:53-66: new transmission, old symptoms
    a.inf_var[slot] = (uint8_t)var;
    a.pvar[p] = (uint8_t)var;
    trans_next[p] = trans_cur[p] = (D.trans_type == JB_TRANS_CONSTANT) ? par[0] : 0.0;
    return;
  }
  a.inf_person[slot] = (int32_t)p;
  a.inf_start[slot] = a.sp->now;
  a.inf_next[slot] = nst > 1 ? tstage[0] : 1e300;  // == traj_time[1]
  a.inf_stage[slot] = 0;
  a.inf_nst[slot] = (uint8_t)nst;
  a.inf_var[slot] = (uint8_t)var;
  a.inf_tag[slot] = (int8_t)D.tag_exposed;
  a.inf_maxtag[slot] = (int8_t)idx;
  a.inf_region[slot] = (uint8_t)region_p;
  a.inf_hosp[slot] = hosp_p;
  a.inf_med[slot] = med_p;
  a.inf_sev[slot] = 0;
  a.pslot[p] = (int32_t)slot;
  a.pvar[p] = (uint8_t)var;
  a.ptag[p] = (int8_t)D.tag_exposed;
  trans_next[p] = (D.trans_type == JB_TRANS_CONSTANT) ? par[0] : 0.0;
  if (!a.fuse_health) trans_cur[p] = trans_next[p];
  // immunity.add_immunity(infection.immunity_ids()) (infection_selector.py:160-162)
  // (cross immunity: Infection.immunity_ids, infection.py:136-191)
  const uint32_t cross = D.cross[var];
  for (int v = 0; v < a.V; ++v)
    if ((cross >> v) & 1u) a.susc[(size_t)v * a.NP + p] = 0.0;
  a.pstate[p] = (uint8_t)(((cross & 1u) ? (st & ~JB_PST_CODE_MASK) : st) | JB_PS_INFECTED);  // susceptibility code of variant 0: 0.0
  atomicAdd(&a.summary[region_p * JB_N_SUMMARY + SUM_NEW_INFECTED], 1u);
  if (a.fuse_health) {
    // update_health_status also covers the infections created in this step
    // (epidemiology.py:217-303 runs after infect_people); k_health handles the older slots
    // Common case decided from registers: still in its first stage, not yet infectious, not in a
    // ward -> the update changes nothing but the "currently infected" counts (same result as the
    // call below, without walking its code and re-reading the slot).
    const double t_inf = a.sp->t_now - a.sp->now;
    const bool quiet = med_p < 0 && nst > 1 && !(t_inf > tstage[0]) &&
                       ((D.trans_type == JB_TRANS_GAMMA && t_inf < par[1]) ||
                        (D.trans_type == JB_TRANS_XNEXP && !(t_inf > par[2])) || D.trans_type == JB_TRANS_CONSTANT);
    if (quiet) {
      if (D.trans_type == JB_TRANS_GAMMA) trans_next[p] = par[3] * 0.0;  // norm * gamma_pdf(x < loc)
      atomicAdd(&a.counters[CNT_INFECTED], 1u);
      atomicAdd(&a.summary[region_p * JB_N_SUMMARY + SUM_CUR_INFECTED], 1u);
    } else {
      HealthAcc acc = {0u, 0u, 0u};
      jb_health_slot(a.h, trans_next, slot, a.sp->t_now, a.sp->flags, a.summary, acc);
      if (acc.inf) atomicAdd(&a.counters[CNT_INFECTED], acc.inf);
      if (acc.hosp) atomicAdd(&a.counters[CNT_HOSP], acc.hosp);
      if (acc.icu) atomicAdd(&a.counters[CNT_ICU], acc.icu);
    }
  }
}

// Which trajectory a new infection follows: max_severity -> outcome slot (symptoms.py:47,110-120;
// health_index.py:151-182), np.searchsorted(side="left") over the person's cumulative outcome probabilities, after
// HealthIndexGenerator.apply_effective_multiplier (health_index.py:183-204) when the person's multiplier is not 1:
// the mass of the outcomes from "severe" upwards (+ the residual) is scaled, the milder ones renormalised; the sums
// run in index order, like the sequential reference.  One lane does it all (the variant of the 16-lane code below).
__device__ __forceinline__ int jb_newinf_outcome(const DiseaseDev& D, const double* cum, double mult, double max_sev) {
  int idx = 0;
  if (mult == 1.0) {
    for (int j = 0; j < D.n_tags; ++j) idx += cum[j] < max_sev ? 1 : 0;
  } else {
    const int max_mild = __ffs(D.severe_mask) - 1 - 2;
    double total = 0.0, mild = 0.0, severe = 0.0, prev = 0.0;
    for (int j = 0; j < D.n_tags; ++j) {
      const double c = cum[j], x = c - (j ? prev : 0.0);
      prev = c;
      total += x;
      if (j < max_mild) mild += x; else severe += x;
    }
    severe += 1 - total;
    const double mod_severe = severe * mult, mod_mild = 1.0 - mod_severe;
    double run = 0.0;
    prev = 0.0;
    for (int j = 0; j < D.n_tags; ++j) {
      const double c = cum[j], pr = c - (j ? prev : 0.0);
      prev = c;
      run += j < max_mild ? pr * mod_mild / mild : pr * mod_severe / severe;
      idx += run < max_sev ? 1 : 0;
    }
  }
  return idx >= D.n_tags ? D.n_tags - 1 : idx;  // residual mass of a cumsum ending at 1-eps
}

// The distribution sample slot j (1..15) of a new infection draws from, or null: 1..8 stage j-1 of trajectory k,
// 9..14 transmission parameter j-9, 15 the second `alpha()` call of the xnexp profile.
__device__ __forceinline__ const JbDist* jb_newinf_dist(const DiseaseDev& D, int j, int k, int nst, uint32_t var) {
  if (j >= 1 && j <= 8) return (j - 1 < nst) ? &D.traj_stage_time[k][j - 1] : nullptr;
  if (j >= 9) {
    const int q = (j == 15) ? 3 : j - 9;
    const int need = D.trans_type == JB_TRANS_GAMMA ? 4 : D.trans_type == JB_TRANS_XNEXP ? 5 : 1;
    if (q < need && (j != 15 || D.trans_type == JB_TRANS_XNEXP)) return &D.trans_par_v[var][q];
  }
  return nullptr;
}

// Sample-major form: a CTA of 16 warps takes 32 new infections at a time; LANE = infection, WARP = sample slot, so a
// warp draws the same quantity for 32 different infections (lanes differ in parameters, seldom in the sampler they
// run) and the sixteen samples of an infection are drawn side by side.  Every warp repeats the short prologue (person
// facts, max_severity, outcome, trajectory: their loads hit L1 after the first warp) instead of waiting for a
// broadcast behind a barrier; the samples meet in shared memory and warp 0 -- whose own slot, max_severity, is part of
// the prologue -- takes the infection slots and the assembling lane's person facts while the others sample, then
// assembles.  Same Philox streams as ever: (person, step, JB_STREAM_SAMPLE0 + slot).
#define JB_NEWINF_THREADS 512
#define JB_NEWINF_GRID (148 * 2)  // two resident CTAs per SM; a CTA walks the chunks blockIdx, blockIdx + grid, ...
__global__ void __launch_bounds__(JB_NEWINF_THREADS, 2) k_newinf(const __grid_constant__ NewInfArgs a, const __grid_constant__ DiseaseDev D) {
  jb_stamp(6);
  __shared__ double s_x[16][32];
  const uint32_t first = a.first_dev ? min(*a.first_dev, a.cap) : 0u;
  const uint32_t n = a.count_dev ? min(*a.count_dev, a.cap - first) : a.count_host;
  const int lane = threadIdx.x & 31, j = threadIdx.x >> 5;
  const uint32_t step_flags = a.sp->flags;
  for (uint32_t i0 = blockIdx.x * 32u; i0 < n; i0 += gridDim.x * 32u) {
    const uint32_t i = i0 + lane;
    bool live = i < n;
    uint32_t p = 0, var = 0, st = 0, gid = 0, region_p = 0, slot = 0;
    int32_t med_p = -1, hosp_p = -1;
    int idx = 0, k = 0, nst = 0;
    if (live) {
      p = a.member[first + i];
      var = a.mutate ? a.mutate_variant : a.variant ? a.variant[first + i] : 0u;
      if (p >= a.N) {  // halo person: tell the home domain (epidemiology.py:359-401)
        if (j == 0) {
          if (a.halo_ret) a.halo_ret[p - a.N] = (uint8_t)(var + 1);
          atomicAdd(&a.counters[CNT_HALO_INF], 1u);
        }
        live = false;
      } else if (step_flags & JB_STEP_NO_NEW_INFECTIONS) {
        live = false;  // only route halo infections (the host creates the local ones)
      }
    }
    if (live) {
      // one round trip for everything indexed by the person
      st = a.pstate[p];
      const uint32_t ext = a.int2ext[p];
      const uint32_t age_raw = a.age[p], sex_p = a.sex[p], ch_p = a.ch[p];
      uint32_t sa_p = 0;
      if (j == 0) { region_p = a.pregion[p]; sa_p = a.psa[p]; med_p = a.pmed[p]; }
      if (!a.mutate && (st & (JB_PS_INFECTED | JB_PS_DEAD))) live = false;  // cannot happen within a step; guards seeding
      if (a.mutate && !(st & JB_PS_INFECTED)) live = false;
      if (live) {
        if (j == 0) {
          slot = a.mutate ? (uint32_t)a.pslot[p] : atomicAdd(&a.counters[CNT_SLOTS_HW], 1u);  // in flight while the others sample
          hosp_p = a.sa_hospital[sa_p];
        }
        gid = a.gid_base + ext;
        if (j <= 8) {  // the trajectory decides the stage-time distributions (slots 1..8) and the assembly (warp 0)
          // symptoms: max_severity -> outcome slot (symptoms.py:47,110-120; health_index.py:151-182)
          JbStream r0;
          jb_stream_init(r0, a.sp->seed, a.sp->step, gid, JB_STREAM_SAMPLE0);
          const double max_sev = jb_stream_next(r0);
          const int age = age_raw > 99u ? 99 : (int)age_raw;
          const int pop = (ch_p && age >= 50) ? 1 : 0;
          const double* cum = a.hi_cum + ((size_t)(pop * 2 + sex_p) * 100 + age) * D.n_tags;
          const double mult = a.effmult ? a.effmult[(size_t)var * a.N + p] : 1.0;
          idx = jb_newinf_outcome(D, cum, mult, max_sev);
          for (int t = 0; t < D.n_traj; ++t)
            if (D.traj_max_tag[t] == idx) k = t;
          nst = D.traj_n_stages[k];
        }
      }
    }
    double x = 0.0;
    if (live && j) {
      const JbDist* dist = jb_newinf_dist(D, j, k, nst, var);
      if (dist) {
        JbStream rs;
        jb_stream_init(rs, a.sp->seed, a.sp->step, gid, JB_STREAM_SAMPLE0 + (uint32_t)j);
        x = jb_sample_dist(rs, dist->type, dist->p);
      }
    }
    s_x[j][lane] = x;
    __syncthreads();
    if (j == 0 && live) {
      if (slot >= a.C) {
        atomicAdd(&a.counters[CNT_OVERFLOW], 1u);
      } else {
        double tstage[JB_MAX_STAGES], tp[7];
#pragma unroll
        for (int s = 0; s < JB_MAX_STAGES; ++s) tstage[s] = s_x[1 + s][lane];
#pragma unroll
        for (int q = 0; q < 7; ++q) tp[q] = s_x[9 + q][lane];
        double* const trans_next = jb_trans_write(a.trans, a.NP, a.sp);
        double* const trans_cur = a.trans + (size_t)(a.sp->tpar & 1u) * a.NP;
        jb_newinf_assemble(a, D, p, var, st, slot, region_p, hosp_p, med_p, idx, k, nst, tstage, tp, trans_next, trans_cur);
      }
    }
    __syncthreads();  // s_x is reused by the next chunk
  }
  jb_stamp_max(3);
}

// Row 4 of the gamma profile's parameters (jb_gamma_front) for uploaded infections [s0, s0 + n): derived on the device
// by the expression the creating kernel uses, so that a restored run continues bit for bit.
__global__ void k_gamma_front(double* __restrict__ inf_par, uint32_t C, uint32_t s0, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = s0 + i;
  inf_par[(size_t)4 * C + s] = jb_gamma_front(inf_par[(size_t)3 * C + s], inf_par[s], inf_par[(size_t)2 * C + s]);
}

// jb_upload_infections, per-person side of the uploaded rows [0, n) -> slots [hw, hw + n): what the creating kernel
// writes for an infection it makes itself (state byte, slot, tag, variant, both halves of the transmission
// probability, cross immunity), plus the slot's mirror of the person's ward.
struct JbCrossMasks { uint32_t c[JB_MAX_VARIANTS]; };
__global__ void k_upload_person_side(const int32_t* __restrict__ person, const int8_t* __restrict__ tag, const uint8_t* __restrict__ var,
                                     const double* __restrict__ prob, uint32_t n, uint32_t hw, uint32_t NP, int V,
                                     const JbCrossMasks cross, int severe_mask, uint8_t* __restrict__ pstate,
                                     int32_t* __restrict__ pslot, int8_t* __restrict__ ptag, uint8_t* __restrict__ pvar,
                                     double* __restrict__ trans0, double* __restrict__ trans1, double* __restrict__ susc,
                                     const int32_t* __restrict__ pmed, int32_t* __restrict__ inf_med) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t p = (uint32_t)person[i];
  const uint32_t vr = var[i];
  const uint32_t cross_r = cross.c[vr < JB_MAX_VARIANTS ? vr : 0];
  uint32_t st = pstate[p];
  st = ((cross_r & 1u) ? (st & ~JB_PST_CODE_MASK) : st) | JB_PS_INFECTED;  // susceptibility to variant 0: 0.0 from now on
  if (JB_TAG_BIT(tag[i]) & severe_mask) st |= JB_PS_SEVERE;
  pstate[p] = (uint8_t)st;
  pslot[p] = (int32_t)(hw + i);
  ptag[p] = tag[i];
  pvar[p] = (uint8_t)vr;
  trans0[p] = prob[i];
  trans1[p] = prob[i];
  for (int v = 0; v < V; ++v)
    if ((cross_r >> v) & 1u) susc[(size_t)v * NP + p] = 0.0;
  inf_med[hw + i] = pmed[p];
}

// Mutation event, selection pass (event/mutation.py:53-58): one thread per infection slot; the chosen people
// are appended to the new-infection list, which k_newinf then walks in mutation mode.
__global__ void __launch_bounds__(256) k_mutate_select(const int32_t* __restrict__ inf_person, const uint8_t* __restrict__ inf_region,
                                                       uint32_t n_slots, const double* __restrict__ prob, const uint32_t* __restrict__ int2ext,
                                                       uint32_t gid_base, uint64_t seed, uint32_t step, uint32_t* __restrict__ member,
                                                       uint32_t cap, uint32_t* __restrict__ count) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  const int32_t p = inf_person[s];
  if (p < 0) return;
  const JbPhilox4 r = jb_philox4x32_10(gid_base + int2ext[p], step, 7u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
  if (jb_u52(r.x[0], r.x[1]) < prob[inf_region[s]]) {
    const uint32_t i = atomicAdd(count, 1u);
    if (i < cap) member[i] = (uint32_t)p;
  }
}

// ------------------------------------------------------------------------------------------
// Halo exchange kernels (the device form of MovablePeople, mpi_wrapper.py:145-293).
// Record: 8 bytes header {hot byte, 7 pad} + max(1,V) doubles
// (infected: [0] = transmission probability; else: susceptibility per variant).
// ------------------------------------------------------------------------------------------
__global__ void k_halo_pack(const uint32_t* __restrict__ out_person, uint32_t n_out, const uint8_t* __restrict__ phot,
                            const double* __restrict__ susc,
                            const double* __restrict__ trans2, uint32_t NP, int V, unsigned char* __restrict__ send, const StepDev* sp) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const double* trans = jb_trans_read(trans2, NP, sp);
  uint32_t p = out_person[i];
  size_t rec = 8 + (size_t)8 * V;
  unsigned char* r = send + (size_t)i * rec;
  uint32_t st = phot[p];  // activity | INFECTED | susceptibility code | variant
  *(uint64_t*)r = (uint64_t)st;
  double* d = (double*)(r + 8);
  if (st & JB_PS_INFECTED) {
    d[0] = trans[p];
    for (int v = 1; v < V; ++v) d[v] = 0.0;
  } else {
    for (int v = 0; v < V; ++v) d[v] = susc[(size_t)v * NP + p];
  }
}

// jb_set_halo_gids: the global id of every visitor that owns a slot in a mirrored supergroup.
__global__ void k_mirror_halo_gids(const uint32_t* __restrict__ halo_mslot, const uint32_t* __restrict__ gids, uint32_t H,
                                   uint32_t* __restrict__ mgid) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H) return;
  const uint32_t slot = halo_mslot[j];
  if (slot != JB_MB_NONE) mgid[slot] = gids[j];
}

__global__ void k_halo_unpack(const unsigned char* __restrict__ recv, uint32_t H, uint32_t N, uint32_t NP, int V,
                              uint8_t* __restrict__ phot, double* __restrict__ susc,
                              double* __restrict__ trans2, uint8_t* __restrict__ halo_ret, const StepDev* sp,
                              const uint32_t* __restrict__ pmslot, uint8_t* __restrict__ marr) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H) return;
  double* trans = trans2 + (size_t)(sp->tpar & 1u) * NP;
  size_t rec = 8 + (size_t)8 * V;
  const unsigned char* r = recv + (size_t)j * rec;
  uint64_t hdr = *(const uint64_t*)r;
  const double* d = (const double*)(r + 8);
  uint32_t p = N + j;
  uint32_t st = (uint32_t)(hdr & 0xFFu);
  phot[p] = (uint8_t)st;
  if (marr) {  // a visitor registered in a mirrored supergroup: its byte goes to its slot as well
    const uint32_t slot = pmslot[p];
    if (slot != JB_MB_NONE) marr[slot] = (uint8_t)jb_hot_to_mirror(st, JB_ACT_PRIMARY);
  }
  if (st & JB_PS_INFECTED) {
    trans[p] = d[0];
  } else {
    for (int v = 0; v < V; ++v) susc[(size_t)v * NP + p] = d[v];
  }
  if (halo_ret) halo_ret[j] = 0;
}

// ---- the same exchange over peer memory ------------------------------------------------------------
// k_halo_push: like k_halo_pack, but every record is stored straight into the destination GPU's window
// (NVLink P2P store); the last block to finish publishes this rank's arrival flag on every peer.
__device__ __forceinline__ void jb_p2p_signal(uint32_t* const* flags, const P2PDev& P, uint32_t* done, uint32_t epoch) {
  __threadfence_system();  // this block's remote stores are visible system-wide before the ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t ticket = atomicAdd(done, 1u);
    if (ticket == gridDim.x - 1) {
      *done = 0;
      __threadfence_system();
      for (int q = 0; q < P.ws; ++q)
        if (q != P.rank) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags[q] + P.rank), "r"(epoch) : "memory");
    }
  }
}

__global__ void __launch_bounds__(256) k_halo_push(const __grid_constant__ P2PDev P, const uint32_t* __restrict__ out_person,
                                                  uint32_t n_out, const uint8_t* __restrict__ phot,
                                                  const double* __restrict__ susc, const double* __restrict__ trans2,
                                                  uint32_t NP, int V, const StepDev* sp, uint32_t* done) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const double* trans = jb_trans_read(trans2, NP, sp);
  if (i < n_out) {
    int q = 0;
    while (q + 1 < P.ws && i >= P.out_off[q + 1]) ++q;
    const uint32_t p = out_person[i];
    const size_t rec = 8 + (size_t)8 * V;
    unsigned char* r = P.recv[q] + ((size_t)P.peer_recv_base[q] + (i - P.out_off[q])) * rec;
    const uint32_t st = phot[p];
    *(uint64_t*)r = (uint64_t)st;
    double* d = (double*)(r + 8);
    if (st & JB_PS_INFECTED) {
      d[0] = trans[p];
      for (int v = 1; v < V; ++v) d[v] = 0.0;
    } else {
      for (int v = 0; v < V; ++v) d[v] = susc[(size_t)v * NP + p];
    }
  }
  jb_p2p_signal(P.flag_people, P, done, sp->epoch);
}

// One thread per peer spins (bounded) until that peer's flag for this step has arrived.
// `dead` is sticky: after one timeout (a peer is gone or out of step) later waits return at once,
// every step reports JB_ECAP, and a queued run drains instead of waiting again and again.
// `cnt_off`: per-rank record offsets of the direction waited for (inbound persons, or outbound persons for the
// return bytes); a rank that exchanges nothing with this one is not waited for (it may not run the exchange at all).
__global__ void k_halo_wait(const uint32_t* flags_self, const __grid_constant__ P2PDev P, int ret, const StepDev* sp,
                            uint32_t* counters, uint32_t* dead) {
  const int q = threadIdx.x;
  const int ws = P.ws, rank = P.rank;
  if (q >= ws || q == rank) return;
  const uint32_t* off = ret ? P.out_off : P.in_off;
  if (off[q + 1] == off[q]) return;
  if (*(volatile uint32_t*)dead) { atomicAdd(&counters[CNT_OVERFLOW], 1u); return; }
  const uint32_t epoch = sp->epoch;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags_self + q) : "memory");
    if ((int32_t)(v - epoch) >= 0) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 10000000000ull) { *dead = 1u; atomicAdd(&counters[CNT_OVERFLOW], 1u); break; }  // 10 s
    __nanosleep(200);
  }
}

// Bytes of the halo persons infected here go to their home GPU's return window.
__global__ void __launch_bounds__(256) k_halo_ret_push(const __grid_constant__ P2PDev P, const uint8_t* __restrict__ ret_local,
                                                      uint32_t H, const StepDev* sp, uint32_t* done) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < H) {
    int s = 0;
    while (s + 1 < P.ws && j >= P.in_off[s + 1]) ++s;
    P.ret[s][(size_t)P.peer_ret_base[s] + (j - P.in_off[s])] = ret_local[j];
  }
  jb_p2p_signal(P.flag_ret, P, done, sp->epoch);
}

// Returned bytes (variant+1 per outbound slot) -> list of residents to infect.
// They are appended BEHIND the step's own new infections (`first` = their count, final by now), which stay readable
// (jb_fetch_new_infections).
__global__ void k_halo_return(const uint8_t* __restrict__ ret, const uint32_t* __restrict__ out_person, uint32_t n_out,
                              uint32_t* __restrict__ member, uint8_t* __restrict__ variant, uint32_t cap,
                              const uint32_t* __restrict__ first, uint32_t* __restrict__ counter) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  uint32_t v = ret[i];
  if (!v) return;
  uint32_t k = min(*first, cap) + atomicAdd(counter, 1u);
  if (k < cap) { member[k] = out_person[i]; variant[k] = (uint8_t)(v - 1); }
}

// ------------------------------------------------------------------------------------------
// Vaccination: VaccinationCampaigns.apply + VaccineTrajectory.update_vaccine_effect for every living
// person, once per simulated day (vaccination_campaign.py:340-385, vaccines.py:109-150,319-376;
// called from epidemiology.py:295-303).  One thread per person; all state is per-person SoA.
// ------------------------------------------------------------------------------------------
struct VaccArgs {
  const VaccDev* tab;
  uint32_t N, NP;
  int V;
  int32_t today;
  uint64_t seed;
  uint32_t gid_base;
  const uint32_t* int2ext;
  uint8_t* pstate;
  const uint8_t* age; const uint8_t* sex; const uint8_t* ch; const uint8_t* pattr;
  double* susc;
  int8_t* vac_dose; int8_t* vac_type; int8_t* traj_camp; int32_t* traj_day0; uint8_t* traj_stage;
  double* prior_eff_inf; double* prior_eff_sym; double* effmult;
  const double* inject;  // per-person uniform (test mode) or null
};

// Dose.get_efficacy (vaccines.py:109-150), t = whole days since the dose was administered.
__device__ __forceinline__ double jb_dose_efficacy(double eff, double prior, double wf, int t, int A, int W, int Z) {
  double pr, fin;
  int n_days, duration;
  if (t > A + W + Z) return wf * eff;
  else if (t > A + W) { pr = eff; fin = wf * eff; n_days = t - (A + W); duration = Z; }
  else if (t > A) return eff;
  else { pr = prior; fin = eff; n_days = t; duration = A; }
  const double m = (fin - pr) / (double)duration;
  return m * (double)n_days + pr;
}

__global__ void __launch_bounds__(256) k_vaccinate(const __grid_constant__ VaccArgs a) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  const uint32_t st = a.pstate[p];
  if (st & JB_PS_DEAD) return;  // also the padding pseudo-persons
  const VaccDev& T = *a.tab;
  const int V = a.V;
  const uint32_t N = a.N;
  const int age_raw = a.age[p], age = min(age_raw, 99);
  const uint32_t attr = a.pattr ? a.pattr[p] : 0u;
  // ---- VaccinationCampaigns.apply ---------------------------------------------------------------
  double prob[JB_MAX_CAMPAIGNS], norm = 0.0;
  int cand[JB_MAX_CAMPAIGNS], n = 0;
  const int dose_now = a.vac_dose[p], type_now = a.vac_type[p];
  for (uint32_t k = 0; k < T.n_campaigns; ++k) {
    const JbVaccinationCampaign& c = T.campaigns[k];
    if (!(c.start_day <= a.today && a.today < c.end_day)) continue;
    const int sd = c.dose_numbers[0];
    bool ok = true;  // has_right_dosage (:182-202)
    if (dose_now >= 0 && sd == 0) ok = false;
    if (sd > 0) {
      if (dose_now < 0 || dose_now != sd - 1) ok = false;
      else if (c.last_dose_vaccines && !((c.last_dose_vaccines >> type_now) & 1u)) ok = false;
    }
    if (!ok) continue;
    bool target;  // is_target_group (:151-180)
    switch (c.group_by) {
      case JB_VG_AGE: target = c.lo <= age_raw && age_raw < c.hi; break;
      case JB_VG_RESIDENCE: target = (a.ch[p] != 0) == (c.lo != 0); break;
      case JB_VG_PRIMARY: target = (int)JB_PA_KIND(attr) == c.lo; break;
      default: target = (int)a.sex[p] == c.lo; break;
    }
    if (!target) continue;
    const int total_days = c.end_day - c.start_day, days_passed = a.today - c.start_day;
    prob[n] = c.coverage * (1 / ((double)total_days - (double)days_passed * c.coverage));  // :248-262
    norm += prob[n];
    cand[n++] = (int)k;
  }
  if (n) {
    const JbPhilox4 r = jb_philox4x32_10(a.gid_base + a.int2ext[p], (uint32_t)a.today, JB_STREAM_VACCINE, 0u,
                                         (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    const double u = a.inject ? a.inject[p] : jb_u52(r.x[0], r.x[1]);
    if (u < norm) {
      const double u2 = jb_u52(r.x[2], r.x[3]);
      int pick = cand[n - 1];
      double c = 0.0;
      for (int i = 0; i < n; ++i) { c += prob[i] / norm; if (c > u2) { pick = cand[i]; break; } }
      // VaccinationCampaign.vaccinate -> Vaccine.generate_trajectory (vaccines.py:515-617)
      for (int v = 0; v < V; ++v) {
        a.prior_eff_inf[(size_t)v * N + p] = 1.0 - a.susc[(size_t)v * a.NP + p];
        a.prior_eff_sym[(size_t)v * N + p] = 1.0 - a.effmult[(size_t)v * N + p];
      }
      a.traj_camp[p] = (int8_t)pick; a.traj_day0[p] = a.today; a.traj_stage[p] = 0;
      a.vac_dose[p] = (int8_t)T.campaigns[pick].dose_numbers[0];
      a.vac_type[p] = (int8_t)T.campaigns[pick].vaccine;
    }
  }
  // ---- VaccineTrajectory.update_vaccine_effect --------------------------------------------------
  const int tc = a.traj_camp[p];
  if (tc < 0) return;
  const JbVaccinationCampaign& c = T.campaigns[tc];
  const JbVaccine& vac = T.vaccines[c.vaccine];
  int admin[JB_MAX_DOSES], day = a.traj_day0[p];
#pragma unroll
  for (int i = 0; i < JB_MAX_DOSES; ++i) { if (i < c.n_doses) day += c.days_to_next_dose[i]; admin[i] = day; }
  const int last = c.n_doses - 1, nl = c.dose_numbers[last];
  if (a.today > admin[last] + vac.days_administered_to_effective[nl] + vac.days_effective_to_waning[nl] + vac.days_waning[nl]) {
    a.traj_camp[p] = -1;  // is_finished: person.vaccine_trajectory = None
    return;
  }
  int stage = a.traj_stage[p];
  const int number_before = c.dose_numbers[stage];
  if (stage < last && a.today >= admin[stage + 1]) ++stage;  // update_trajectory_stage
  a.traj_stage[p] = (uint8_t)stage;
  const int num = c.dose_numbers[stage];
  const int t = a.today - admin[stage];
  for (int v = 0; v < V; ++v) {
#pragma unroll
    for (int kind = 0; kind < 2; ++kind) {
      const double* tab = kind == 0 ? vac.sterilisation : vac.symptomatic;
      const double eff = tab[((size_t)num * V + v) * 100 + age];
      double* prior_arr = kind == 0 ? a.prior_eff_inf : a.prior_eff_sym;
      const double prior0 = prior_arr[(size_t)v * N + p];
      const double prior = stage == 0 ? prior0 : tab[((size_t)c.dose_numbers[stage - 1] * V + v) * 100 + age] * vac.waning_factor;
      const double e = jb_dose_efficacy(eff, prior, vac.waning_factor, t, vac.days_administered_to_effective[num],
                                        vac.days_effective_to_waning[num], vac.days_waning[num]);
      const double updated = 1.0 - e, prior_v = 1.0 - prior0;
      const double out = prior_v < updated ? prior_v : updated;
      if (kind == 0) {
        a.susc[(size_t)v * a.NP + p] = out;
        if (v == 0) {  // the 2-bit code the interaction kernels read instead of the f64
          const uint32_t code = out == 0.0 ? JB_SC_ZERO : out == 1.0 ? JB_SC_ONE : JB_SC_READ;
          a.pstate[p] = (uint8_t)((st & ~JB_PST_CODE_MASK) | code);
        }
      } else {
        a.effmult[(size_t)v * N + p] = out;
      }
    }
  }
  if (num != number_before) a.vac_dose[p] = (int8_t)num;
}

// First node of a step: block 0 pulls the step's parameter block out of the pinned host ring
// (zero-copy reads over PCIe; the slot is chosen by a device-side step count that advances with every
// execution, mirroring the host's), all blocks zero the per-step counters.  One kernel node instead of
// a DMA copy and a memset node in front of the placement.
__global__ void __launch_bounds__(256) k_step_begin(const uint32_t* host_ring, uint32_t slot_words, uint32_t* __restrict__ d_step,
                                                    uint32_t n_words, uint32_t* __restrict__ seq,
                                                    uint32_t* __restrict__ zero_base, uint32_t n_zero) {
  jb_stamp(0);
  if (blockIdx.x == 0) {
    const uint32_t sq = *seq;
    const volatile uint32_t* src = host_ring + (size_t)(sq & 3u) * slot_words;
    for (uint32_t i = threadIdx.x; i < n_words; i += blockDim.x) d_step[i] = src[i];
    __syncthreads();
    if (threadIdx.x == 0) *seq = sq + 1u;
  }
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_zero; i += gridDim.x * blockDim.x) zero_base[i] = 0u;
}

// Last node of a step: the counter block goes straight into pinned host memory (posted writes; visible
// to the host once the stream has been synchronised).
__global__ void __launch_bounds__(256) k_step_end(const uint32_t* __restrict__ counters, uint32_t n, uint32_t* host_out) {
  jb_stamp(7);
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) host_out[i] = counters[i];
}

// Recompute the susceptibility codes kept in `pstate` after the host replaced susceptibilities.
__global__ void k_refresh_codes(uint8_t* __restrict__ pstate, const double* __restrict__ susc, uint32_t N) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double s = susc[i];
  const uint32_t code = s == 0.0 ? JB_SC_ZERO : s == 1.0 ? JB_SC_ONE : JB_SC_READ;
  pstate[i] = (uint8_t)((pstate[i] & ~JB_PST_CODE_MASK) | code);
}

// Residence group of every resident (the location of a death outside hospital), one thread per group.
__global__ void k_res_group(const uint32_t* __restrict__ grp_off, const uint32_t* __restrict__ reg_member,
                            const uint16_t* __restrict__ reg_info, uint32_t G, uint32_t N, int32_t* __restrict__ pres_group) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  for (uint32_t m = grp_off[g]; m < grp_off[g + 1]; ++m)
    if (JB_INFO_ACT(reg_info[m]) == JB_ACT_RESIDENCE && reg_member[m] < N) pres_group[reg_member[m]] = (int32_t)g;
}

__global__ void k_fill_f64(double* x, size_t n, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}
