// Internal declarations shared by the translation units of libjune_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/june_b200.h"

// ---- device counters (uint32 array `counters`) -------------------------------------------
// The counter block is laid out as [persistent | per-step | per-region summary | dynamic-bin
// counts]; everything from CNT_STEP_BEGIN to the end is zeroed by ONE memset at step start.
enum {
  CNT_SLOTS_HW = 0,    // infection slot high-water mark (persistent)
  CNT_DEAD_TOTAL,      // persistent
  CNT_SLOTS_PREV,      // high-water mark at the start of the step (snapshot by k_place)
  CNT_STEP_BEGIN,
  CNT_NEWINF = CNT_STEP_BEGIN,  // new infections appended this step (local + halo)
  CNT_NOACT,           // persons without any activity this step
  CNT_DRAWS,           // Bernoulli draws consumed this step
  CNT_DYN,             // (unused)
  CNT_RECOVERED_STEP,
  CNT_DEATHS_STEP,
  CNT_HALO_INF,
  CNT_OVERFLOW,        // a capacity was exceeded
  CNT_HEV,             // health events appended this step (JB_STEP_RECORD_EVENTS)
  CNT_INFECTED,        // currently infected / in hospital / in icu (recomputed by k_health)
  CNT_HOSP,
  CNT_ICU,
  CNT_ACT0,            // persons placed per activity (5 words: JB_ACT_MEDICAL..JB_ACT_RESIDENCE)
  CNT_ACT4 = CNT_ACT0 + 4,
  CNT_VIS_LIST,        // residence visits: people who chose a leisure destination this step (deferred placement)
  CNT_VIS_N,           // households being visited this step
  CNT_N                // followed by summary[R][JB_N_SUMMARY], then dyn_cnt[n_bins]
};

// per-region summary columns (records_writer.py:151-164)
enum { SUM_NEW_INFECTED = 0, SUM_CUR_INFECTED, SUM_CUR_HOSP, SUM_CUR_ICU, SUM_DEATHS, SUM_HOSP_DEATHS, SUM_RECOVERED, SUM_ADMISSIONS, SUM_ICU_ADMISSIONS };

#define JB_TAG_BIT(tag) (1 << ((tag) + 2))

// ---- internal person bytes -----------------------------------------------------------------
// `pstate` (persistent): bits 0-1 susceptibility code of variant 0, bit 3 INFECTED, 4 DEAD,
// 5 SEVERE, 6 MEDICAL (the JB_PS_* positions of the ABI byte; the activity bits are NOT kept
// here).  `phot` (rebuilt by k_place every step, the only per-person byte the interaction
// kernels read): bits 0-2 activity of the step, bit 3 INFECTED, bits 4-5 susceptibility code,
// bits 6-7 variant of the infection.  The code spares the 8-byte susceptibility gather for the
// common values: 0 -> 0.0 (immune), 1 -> 1.0, 2 -> read `susc`.
#define JB_SC_ZERO 0u
#define JB_SC_ONE 1u
#define JB_SC_READ 2u
#define JB_PST_CODE_MASK 0x03u
#define JB_HOT_CODE(x) (((x) >> 4) & 3u)
#define JB_HOT_VAR(x) ((x) >> 6)

// Mirror byte (slot-ordered copy of the person state for the primary-activity supergroups): the persistent
// pstate bits -- 0-1 susceptibility code, 3 INFECTED, 4 DEAD (= absent), 5 SEVERE, 6 MEDICAL -- plus the
// variant in bits 2 and 7.  Whether the member is present follows from the byte and the step's activity mask /
// flags; a placement with per-person policies writes a per-step copy with bit 4 = "not at this slot".
#define JB_MB_ABSENT 0x10u
#define JB_MB_NONE 0xFFFFFFFFu
__host__ __device__ inline uint32_t jb_mirror_byte(uint32_t pstate, uint32_t var) {
  return (pstate & 0x7Bu) | ((var & 1u) << 2) | ((var & 2u) << 6);
}
// absent if any of these bits is set: dead, in hospital (when medical facilities are open), severe + stay-home flag
__host__ __device__ inline uint32_t jb_mirror_absmask(uint32_t activity_mask, uint32_t flags) {
  return JB_MB_ABSENT | ((activity_mask & (1u << JB_ACT_MEDICAL)) ? (uint32_t)JB_PS_MEDICAL : 0u) |
         ((flags & JB_STEP_SEVERE_STAY_HOME) ? (uint32_t)JB_PS_SEVERE : 0u);
}
// -> the hot-byte format of `phot` (activity | INFECTED | code << 4 | variant << 6)
__host__ __device__ inline uint32_t jb_mirror_to_hot(uint32_t mb, uint32_t absmask, uint32_t act) {
  return ((mb & absmask) ? (uint32_t)JB_ACT_NONE : act) | (mb & 0x08u) | ((mb & 0x03u) << 4) | ((mb & 0x04u) << 4) | (mb & 0x80u);
}
// hot byte of a halo person (as its home domain sent it) -> mirror byte for a slot fed by `act`
__host__ __device__ inline uint32_t jb_hot_to_mirror(uint32_t hot, uint32_t act) {
  return ((hot & 7u) == act ? 0u : JB_MB_ABSENT) | (hot & 0x08u) | ((hot >> 4) & 0x03u) | ((hot >> 4) & 0x04u) | (hot & 0x80u);
}

// tile slot descriptor (uint8): bits 0-2 feeding activity, 3 VALID, 4 FIRST slot of its group,
// 5-6 local subgroup (tile mode needs dim <= 4)
#define JB_TI_VALID 0x08u
#define JB_TI_FIRST 0x10u
#define JB_TI_LSG(x) (((x) >> 5) & 3u)
#define JB_TI_ACT(x) ((x) & 7u)

// Work item kinds of k_span (item.y): bits 28-31 kind, low bits count.
#define JB_TILE_GINFO_MIXED 0xFFFFFFFFu  // tile_ginfo: the tile's groups do not share one group word
#define JB_SPAN_KIND_SMALL 1u  // count = tiles of small groups (<= 16); tile_group0 = ordinal of the tile's first group in span_group
#define JB_SPAN_KIND_CLASS 2u  // count = groups, bits 8-15 = lanes per group (4, 8, 16 or 32: 16 slots per lane); item.z = first ordinal
#define JB_SPAN_KIND_BIG 3u    // one group of more than 512 slots; item.z = its ordinal

// Dynamic-member bins: one bin of `cap` entries per group of every supergroup with has_dynamic.
#define JB_MAX_DYN_RANGES 12
struct DynRanges {
  uint32_t n;
  uint32_t first[JB_MAX_DYN_RANGES], count[JB_MAX_DYN_RANGES];  // group range of the supergroup
  uint32_t cnt_base[JB_MAX_DYN_RANGES];   // index of the range's first counter in dyn_cnt
  uint32_t slot_base[JB_MAX_DYN_RANGES];  // offset of the range's first bin in dyn_bin (entries)
  uint32_t cap[JB_MAX_DYN_RANGES];        // entries per bin (power of two)
};

// Leisure tables on the device (jb_set_leisure / jb_set_leisure_tables).
struct LeisureDev {
  uint32_t K, n_tables;
  const uint32_t* parea;   // [N] internal numbering
  const uint32_t* off;     // [(areas * K) + 1]
  const uint32_t* venue;
  const double* does;      // [tables][R][2][100]
  const double* cum;       // [tables][R][2][100][K]
  const int32_t* inject;   // [N] venue per person (test mode) or null
  uint32_t sg_first[JB_MAX_LEISURE], sg_count[JB_MAX_LEISURE];
  int32_t lsg[JB_MAX_LEISURE];
};

// Active individual policies + regional compliance (jb_set_individual_policies), device resident.
#define JB_MAX_REGIONS_POLICY 64
struct PolicyDev {
  uint32_t n, pad_;
  JbIndividualPolicy pol[JB_MAX_POLICIES];
  double rc[JB_MAX_REGIONS_POLICY];
};

// Residence visits on the device (jb_set_visits).  A step with visits places its leisure-goers in three moves:
// everybody draws a destination (k_place, nothing is scattered yet), the order-dependent part of the reference --
// "a household that is being visited keeps its residents home; the people are walked in world order" -- is resolved
// by a fixed-point iteration over the (few) people who drew one, then they are scattered or sent home.
#define JB_VIS_NONE 0xFFFFFFFFu
#define JB_VIS_DONE (1u << 27)  // `choice`: lsg << 28 | DONE | group (group indices stay below 2^27)
#define JB_VIS_BIN 32          // visitors per household and step (= the smallest bin the group kernel sorts)
#define JB_VIS_ITERATIONS 8    // chains "my household is visited by somebody whose household is visited by ..." longer
                               // than this do not occur (each link has a probability of a few per cent)
struct VisitsDev {
  int32_t activity = -1;       // index among the leisure activities, -1 = no visits
  uint32_t hh_first, hh_count; // groups of the households supergroup
  uint32_t ch_first, ch_count; // groups of the care-home supergroup (count 0: none)
  uint32_t ch_lsg;
  const uint32_t *hh_off, *hh, *ch_off, *ch;
  double p_household;          // P(household | the person's household is linked to both kinds)
  const int32_t* pres_group;   // residence group of every resident
  uint32_t* choice;            // [N] this step's leisure destination (lsg << 28 | group) or JB_VIS_NONE
  uint32_t* list;              // [N] the people with a destination (count: CNT_VIS_LIST)
  uint32_t* hv[3];             // [households] smallest caller's index of an EXECUTED visitor (rotating buffers)
  uint32_t* vslot;             // [households] index in the visited list, or JB_VIS_NONE
  uint32_t *vis_group, *vis_cnt, *vis_bin;  // visited list: group, visitors, their (lsg << 28 | caller's index)
  uint32_t vis_cap;
};

// Vaccines and campaigns on the device (jb_set_vaccination); efficacy tables as device pointers.
struct VaccDev {
  uint32_t n_vaccines, n_campaigns;
  JbVaccine vaccines[JB_MAX_VACCINES];
  JbVaccinationCampaign campaigns[JB_MAX_CAMPAIGNS];
};

// Commute tables on the device (jb_set_commute).
struct CommuteDev {
  const int32_t* person;     // [N] internal numbering: -1 | city << 1 | (station << 1) | 1
  const uint32_t* cs_off;    // city -> stations
  const uint32_t* cs;
  const uint32_t* st_off;    // station -> transports (groups)
  const uint32_t* st;
  const int32_t* inject;     // [N] transport per person (test mode) or null
};

// Per-step scalars, resident on the device so that a captured CUDA graph of the step can be
// replayed: one h2d copy from a pinned ring buffer precedes every step.  The processed beta
// table [n_beta][n_regions] follows the struct in the same allocation.
// Debug timeline (JB_STAMPS=1): every kernel's first thread stamps %globaltimer at entry; printed per
// step shape when the world is destroyed.  IDs: 0 begin, 1 place, 2 tiles, 3 groups (first), 4 groups
// (last), 5 health, 6 newinf, 7 end.
__device__ unsigned long long* g_stamps = nullptr;
__device__ unsigned int g_stamp_seq = 0;
#define JB_STAMP_RING 256
__device__ __forceinline__ unsigned long long jb_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Latest time any block passed a point (id 0: tiles phase A done, 1: tiles done, 2: place done, 3: newinf done).
__device__ __forceinline__ void jb_stamp_max(int id) {
  if (threadIdx.x == 0) {
    unsigned long long* s = g_stamps;
    if (s) atomicMax(s + (size_t)JB_STAMP_RING * 8 + (size_t)(g_stamp_seq % JB_STAMP_RING) * 4 + id, jb_globaltimer());
  }
}
__device__ __forceinline__ void jb_stamp(int id) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long* s = g_stamps;
    if (s) {
      const unsigned long long t = jb_globaltimer();
      unsigned long long* slot = s + (size_t)(g_stamp_seq % JB_STAMP_RING) * 8;
      if (id == 0) {
        const unsigned int q = ++g_stamp_seq;
        slot = s + (size_t)(q % JB_STAMP_RING) * 8;
        for (int k = 1; k < 8; ++k) slot[k] = (k == 4) ? 0ull : ~0ull;
        slot[0] = t;
        unsigned long long* ext = s + (size_t)JB_STAMP_RING * 8 + (size_t)(q % JB_STAMP_RING) * 4;
        ext[0] = ext[1] = ext[2] = ext[3] = 0ull;
      } else if (id == 4) {
        atomicMax(slot + 4, t);
      } else {
        atomicMin(slot + id, t);
      }
    }
  }
}

struct StepDev {
  double now, dt, t_now;
  uint64_t seed;
  uint32_t step, activity_mask, flags, n_inject;
  uint32_t leisure_table, epoch;   // epoch: steps since create (arrival-flag value of the P2P exchange)
  uint32_t tpar, pad_;             // tpar: which half of the double-buffered transmission probabilities this step READS
};

// Peer-memory halo exchange (jb_p2p_connect): windows of all ranks as seen from this GPU.
struct P2PDev {
  int rank, ws;
  unsigned char* recv[JB_MAX_RANKS];   // inbound person records of rank q
  uint8_t* ret[JB_MAX_RANKS];          // inbound "infected abroad" bytes of rank q
  uint32_t* flag_people[JB_MAX_RANKS]; // flag_people[q][r]: records of rank r have arrived at q
  uint32_t* flag_ret[JB_MAX_RANKS];
  uint32_t out_off[JB_MAX_RANKS + 1], in_off[JB_MAX_RANKS + 1];
  uint32_t peer_recv_base[JB_MAX_RANKS], peer_ret_base[JB_MAX_RANKS];
};

// Arguments of the interaction (group) kernels; passed by value as a __grid_constant__.
struct GroupArgs {
  const uint32_t* grp_off;
  const uint32_t* grp_info;
  const uint32_t* grp_aux;
  const uint32_t* reg_member;
  const uint16_t* reg_info;
  const uint8_t* phot;       // per-step person byte (see above)
  const double* susc;
  const double* trans;       // both halves of the double buffer ([2][NP]); the step reads half sp->tpar
  const uint32_t* halo_gid;
  const uint8_t* school_years;
  const double* cm;
  const double* beta;
  const double* beta_scale;
  const uint32_t* dyn_cnt;   // per dynamic group: members scattered this step (or null)
  const uint32_t* dyn_bin;   // [bins][bin_cap] entries (lsg << 28) | person
  uint32_t bin_cap, dyn_bin_base;
  uint32_t g0, g1;
  const uint32_t* group_list; // explicit groups for the warp kernel (big groups), or null = [g0, g1)
  uint32_t n_list;
  const uint32_t* int2ext;    // internal -> caller person index
  const uint32_t* ext2int;    // caller -> internal person index
  const uint32_t* tile_group0;
  const uint4* tile_static;    // per tile: {valid slots, first slots of the groups, subgroup plane 0, subgroup plane 1}, or null
  int uniform_act;             // the activity that feeds every slot of the supergroup (tile_static is only set when there is one)
  const uint32_t* tile_ginfo;  // per tile: the grp_info word its groups share, or JB_TILE_GINFO_MIXED (null: always gather)
  const uint32_t* tp_member;
  const uint8_t* tp_info;
  uint32_t n_tiles, stream_base;
  // mirrored supergroup (null otherwise): slot-ordered state bytes of this step (JB_MB_*), the supergroup's first
  // slot is stream_base; list_slot[i] = first slot (relative to stream_base) of group group_list[i]
  const uint8_t* marr;
  const uint32_t* mgid;      // [mirror slots] global person id of the slot's member (the Philox counter word)
  const uint32_t* list_slot;
  // residence visits: groups with a slot in `vslot` (indexed by group - g0) are left to the visited-list launch;
  // that launch takes its groups from `group_list` with the count on the device, the visitors of list entry i from
  // dyn_bin[i * bin_cap ...] / dyn_cnt[i], and the beta rule `beta_rule - 1` instead of the group's own
  const uint32_t* vslot;
  const uint32_t* n_list_dev;
  int dyn_by_list;
  int beta_rule;
  const uint4* span_items;   // k_span work items (jb_pack_span): {first slot, kind | count, first ordinal, 0}
  const uint32_t* span_group; // groups of the supergroup in slot order
  uint32_t n_span;
  int dim, kind, L;
  uint32_t static_acts;       // activities that feed registered slots of this supergroup (bit per JB_ACT_*)
  uint32_t N, NP;
  int V, R;
  const StepDev* sp;         // dt, seed, step, n_inject
  uint32_t gid_base;
  uint32_t* counters;
  uint32_t* newinf_member;
  uint8_t* newinf_var;
  uint32_t newinf_cap;
  // infection events (JB_STEP_RECORD_EVENTS): same index as newinf_member; null otherwise
  uint32_t* ev_infected;
  uint32_t* ev_infector;
  uint32_t* ev_group;
  uint8_t* ev_var;
  double* lambda;            // null unless JB_STEP_RECORD_LAMBDA
  const double* inject;      // null unless JB_STEP_INJECT_UNIFORMS
  const uint32_t* draw_base; // exclusive scan of draw_cnt (inject mode)
  uint32_t* draw_cnt;
  int mode;                  // 0 = interact, 1 = only count draws per group
};

struct DiseaseDev {
  int n_tags, rec_mask, dead_mask, hosp_mask, icu_mask, severe_mask, tag_exposed, n_traj;
  int traj_max_tag[16];
  int traj_n_stages[16];
  int traj_stage_tag[16][JB_MAX_STAGES];
  JbDist traj_stage_time[16][JB_MAX_STAGES];
  int trans_type;
  JbDist trans_par[6];
  JbDist trans_par_v[JB_MAX_VARIANTS][6];   // per-variant distributions (copies of trans_par unless jb_set_variants)
  uint32_t cross[JB_MAX_VARIANTS];          // bit u of cross[v]: an infection with v protects against u
};

// Tile decomposition of one supergroup (built once at world create).
struct SgTiles {
  int mode = 0;               // 0 = warp per group only, 1 = gather tiles, 2 = stream tiles, 3 = mirror tiles
  bool has_small = false;     // some group is processed by the tile kernel
  uint32_t *list_w_slot = nullptr, *list_c_slot = nullptr;  // mirror: first slot of the listed groups
  uint4* span_items = nullptr;  // mirror, one subgroup per group: work items of k_span (replaces the tile kernel and the warp list)
  uint32_t* span_group = nullptr;
  uint32_t n_span = 0;
  uint32_t n_tiles = 0, stream_base = 0;
  uint32_t* tile_group0 = nullptr;
  uint32_t* tile_ginfo = nullptr;
  uint4* tile_static = nullptr;
  int uniform_act = -1;
  uint32_t* tp_member = nullptr;
  uint8_t* tp_info = nullptr;
  // groups not handled by tiles, longest first: <= 128 members and static -> one warp each,
  // the others -> one CTA each
  uint32_t *list_w = nullptr, *list_c = nullptr;
  uint32_t n_w = 0, n_c = 0;
  double mean_len = 0;          // mean slots per tiled group (picks the lanes-per-group flavour)
};

struct JbWorld {
  int device = 0;
  cudaStream_t stream = nullptr;
  // the interaction kernels of the active supergroups are independent: they run concurrently on
  // side streams forked from / joined into `stream`
  // (up to three independent launches per supergroup: tiles, warp list, CTA list)
  cudaStream_t side[JB_MAX_SUPERGROUPS][3] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[JB_MAX_SUPERGROUPS][3] = {};
  cudaStream_t side_health = nullptr;
  cudaEvent_t ev_fork_health = nullptr, ev_join_health = nullptr;
  // arrival of the visitors (wait + unpack) runs beside the supergroups that hold no visitor registration
  cudaStream_t side_halo = nullptr;
  cudaEvent_t ev_fork_halo = nullptr, ev_halo_ready = nullptr;
  bool halo_async = false;                 // this step: ev_halo_ready gates the supergroups with visitors
  std::vector<uint8_t> sg_has_halo;        // supergroup has registrations of visitors (persons >= N)
  bool serial = false;
  // N = residents in the INTERNAL numbering (includes padding slots of the streamed supergroup),
  // N_ext = residents as the caller numbers them
  uint32_t N = 0, N_ext = 0, H = 0, NP = 0, G = 0, M = 0, V = 1, R = 1, SA = 0, C = 0, newinf_cap = 0;
  std::vector<uint32_t> h_ext2int, h_int2ext;  // int2ext == 0xFFFFFFFF for padding
  uint32_t *int2ext = nullptr, *ext2int = nullptr;
  std::vector<SgTiles> tiles;
  uint32_t n_beta = 0, n_scale = 0, n_out = 0;
  uint64_t id_base = 0;
  std::vector<JbSupergroup> sgs;
  std::vector<int> sg_L;          // max local subgroups per supergroup
  std::vector<uint32_t> sg_static_acts;  // feeding activities of the supergroup's registered slots
  DiseaseDev dis_host;
  // static device arrays
  uint8_t *age = nullptr, *sex = nullptr, *ch = nullptr, *phas = nullptr, *school_years = nullptr;
  uint32_t *psa = nullptr, *grp_off = nullptr, *grp_info = nullptr, *grp_aux = nullptr, *reg_member = nullptr;
  uint16_t *reg_info = nullptr, *sa_region = nullptr;
  int32_t* sa_hospital = nullptr;
  double *cm = nullptr, *beta_scale = nullptr, *hi_cum = nullptr;
  DiseaseDev* dis = nullptr;
  // person state
  uint8_t *pstate = nullptr, *phot = nullptr, *pvar = nullptr;
  // mirror (primary-activity supergroups): persistent slot-ordered state, per-step copy written by the placement
  // with per-person policies, the byte last pushed per person, person -> slot
  uint8_t *mstate = nullptr, *mstep = nullptr, *pmir = nullptr;
  uint32_t* pmslot = nullptr;
  uint32_t* mgid = nullptr;   // [mir_slots] global id of the slot's member
  uint32_t mir_slots = 0;
  int8_t* ptag = nullptr;
  double* susc = nullptr;
  // Transmission probabilities are double buffered: the interaction kernels of a step read trans[trans_par]
  // (written by the previous step's health update) while this step's health update writes trans[trans_par ^ 1]
  // concurrently; the buffers swap when the step finishes.  Only infected people's entries are meaningful.
  double* trans[2] = {nullptr, nullptr};
  uint32_t trans_par = 0;
  int32_t *pmed = nullptr, *pslot = nullptr;
  uint32_t* halo_gid = nullptr;
  uint32_t* out_person = nullptr;
  // infection slots
  int32_t* inf_person = nullptr;
  double *inf_start = nullptr, *inf_par = nullptr, *inf_next = nullptr, *traj_time = nullptr;
  uint8_t *inf_stage = nullptr, *inf_nst = nullptr, *inf_var = nullptr;
  int8_t *inf_tag = nullptr, *inf_maxtag = nullptr, *traj_tag = nullptr;
  uint8_t *inf_region = nullptr, *inf_sev = nullptr;   // slot-resident mirrors of person facts
  int32_t *inf_hosp = nullptr, *inf_med = nullptr;
  std::vector<uint8_t> h_pregion;                       // host copies for jb_upload_infections
  std::vector<int32_t> h_phosp;
  // per step
  uint32_t* counters = nullptr;   // CNT_N words + summary + dyn_cnt (one allocation)
  uint32_t* summary = nullptr;    // [R][JB_N_SUMMARY], inside the counter block
  uint8_t* pregion = nullptr;     // region of every resident
  size_t counter_words = 0;
  uint32_t* newinf_member = nullptr;
  uint8_t* newinf_var = nullptr;
  uint32_t *ev_infected = nullptr, *ev_infector = nullptr, *ev_group = nullptr;  // allocated on first use
  uint8_t* ev_var = nullptr;
  uint32_t ev_count = 0;
  uint32_t *hev_person = nullptr, *hev_kind = nullptr;   // health events (deaths, recoveries, admissions, discharges)
  int32_t* hev_group = nullptr;
  uint32_t hev_cap = 0;
  uint32_t* pol_subjects = nullptr;                      // long-distance commuters (subjects of LimitLongCommute)
  uint32_t n_pol_subjects = 0;
  uint32_t halo_act_mask = 0xFFFFFFFFu;                  // activities under which halo persons are registered anywhere (jb_set_halo_activities)
  bool pol_subject_mode = false;                         // every active individual policy has a static subject list
  int32_t* pres_group = nullptr;                         // residence group of every resident (death location)  // events of the last step that recorded them
  double* beta = nullptr;
  double* lambda = nullptr;
  double* inject = nullptr;
  size_t inject_cap = 0, n_inject = 0;
  uint32_t *draw_cnt = nullptr, *draw_base = nullptr;
  uint32_t *dyn_cnt = nullptr, *dyn_bin = nullptr;
  uint32_t n_bins = 0;
  uint32_t sg_bin_base[JB_MAX_SUPERGROUPS] = {};   // first counter of the supergroup in dyn_cnt
  uint32_t sg_bin_slot[JB_MAX_SUPERGROUPS] = {};   // first entry of the supergroup in dyn_bin
  uint32_t sg_bin_cap[JB_MAX_SUPERGROUPS] = {};    // entries per bin
  DynRanges dyn_ranges = {};
  // leisure
  VisitsDev vz = {};
  int visits_beta = 0, visits_sg = -1;     // beta rule of a visited household; the households' supergroup
  LeisureDev lz = {};
  uint32_t *lz_parea = nullptr, *lz_off = nullptr, *lz_venue = nullptr;
  double *lz_does = nullptr, *lz_cum = nullptr;
  int32_t* lz_inject = nullptr;
  // peer-memory halo exchange
  bool p2p_on = false;
  P2PDev p2p = {};
  unsigned char* p2p_window = nullptr;   // this rank's window (cudaMalloc, IPC-exported)
  std::vector<void*> p2p_opened;         // peers' windows opened with cudaIpcOpenMemHandle
  uint32_t* p2p_done = nullptr;          // last-block tickets of the two push kernels
  uint8_t* p2p_ret_local = nullptr;      // [H] variant+1 of halo persons infected here (before it is pushed home)
  size_t p2p_recv_bytes = 0, p2p_ret_bytes = 0;
  // vaccination
  VaccDev* d_vacc = nullptr;
  std::vector<double*> vacc_tables;
  int8_t *vac_dose = nullptr, *vac_type = nullptr, *traj_camp = nullptr;
  int32_t* traj_day0 = nullptr;
  uint8_t* traj_stage = nullptr;
  double *prior_eff_inf = nullptr, *prior_eff_sym = nullptr, *effmult = nullptr, *vac_inject = nullptr;
  bool vac_inject_on = false;
  uint32_t n_campaigns = 0;
  // commute
  CommuteDev cmt = {};
  int32_t *cm_person = nullptr, *cm_inject = nullptr;
  uint32_t *cm_cs_off = nullptr, *cm_cs = nullptr, *cm_st_off = nullptr, *cm_st = nullptr;
  // individual policies
  PolicyDev* d_pol = nullptr;
  uint32_t n_pol = 0;
  uint8_t* pattr = nullptr;
  uint8_t* pwork_region = nullptr;
  double* pol_inject = nullptr;
  bool any_dynamic = false;
  void* cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  // per-step parameter block: device copy + pinned ring (written by the host, copied h2d per step)
  unsigned char* d_step = nullptr;   // StepDev + beta table
  unsigned char* d_step_aux = nullptr;  // StepDev for jb_infect (seeding outside a step)
  unsigned char* h_step_ring = nullptr;   // pinned, 4 slots of step_stride bytes; read by k_step_begin (zero copy)
  uint32_t* d_step_ring = nullptr;        // the same memory as seen from the device
  size_t step_stride = 0;
  unsigned long long* stamps = nullptr;    // JB_STAMPS=1 debug timeline
  uint32_t* d_step_seq = nullptr;         // device-side count of executed k_step_begin (ring slot = count & 3)
  uint32_t* d_h_counters = nullptr;       // h_counters as seen from the device
  unsigned char* h_step[4] = {};
  cudaEvent_t ev_step_buf[4] = {};
  size_t step_bytes = 0;
  uint64_t step_seq = 0;
  uint32_t* h_counters = nullptr;
  // captured step graphs, keyed by (phase, activity mask, flags, buffers)
  struct GraphEntry { uint64_t k0, k1, k2; cudaGraphExec_t exec; uint64_t launches; };
  std::vector<GraphEntry> graphs;
  bool use_graphs = true;
  // step in flight
  JbStepParams cur;
  bool in_step = false;
  // timing
  int timing = 0;
  bool ev_detail = false;
  cudaEvent_t ev_step0 = nullptr, ev_step1 = nullptr, ev_int0 = nullptr, ev_int1 = nullptr;
  cudaEvent_t ev_sg0[JB_MAX_SUPERGROUPS] = {}, ev_sg1[JB_MAX_SUPERGROUPS] = {};
  bool sg_pending[JB_MAX_SUPERGROUPS] = {};
  double acc_sg_ms[JB_MAX_SUPERGROUPS] = {};
  uint64_t acc_sg_launches[JB_MAX_SUPERGROUPS] = {};
  double acc_int_ms = 0, acc_step_ms = 0;
  uint64_t acc_steps = 0, acc_launches = 0;
  bool ev_pending = false;
  // timing level 1 keeps a ring of event pairs so that timed steps can be queued without a host
  // synchronisation per step; pairs are folded into acc_step_ms when the ring wraps or on read
  std::vector<cudaEvent_t> ring0, ring1;
  uint64_t ring_head = 0, ring_tail = 0;
  unsigned char* flush_buf = nullptr;
  size_t flush_bytes = 0;
  uint64_t last_int_bytes = 0, last_step_bytes = 0;
  uint64_t launches = 0;
};

void jb_set_error(const std::string& msg);
