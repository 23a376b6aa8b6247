/*
 * june_b200.h -- C ABI of the B200-native replacement for JUNE's per-timestep hot path
 * (`Simulator.do_timestep`, reference june/simulator.py:346-626).
 *
 * The reference has no FFI: its only extension point is overriding the Python method
 * `Simulator.do_timestep` (june/simulator.py:192,256,273 are cls-polymorphic; the
 * `ActivityManager` class attribute at :80 swaps the placement stage).  This header is the
 * boundary a maintainer binds from that override with ctypes.  Plain
 * pointers and sizes only; no torch / CUDA types in any signature.  Pointers named `d_*` are
 * device addresses (e.g. `torch.Tensor.data_ptr()`); everything else is host memory owned by
 * the caller and only borrowed for the duration of the call.
 *
 * The very same ABI (prefix `jo_` instead of `jb_`) is exported by the CPU oracle in
 * oracle/june_oracle.c; that library is test infrastructure and is never loaded by the product.
 *
 * Threading: one host thread per world (= per GPU / rank).  All work is enqueued on one
 * library-owned CUDA stream per world; `jb_step` returns after enqueue, `jb_sync`/`jb_fetch_*`
 * wait.  Errors: every function returns 0 on success or a JB_E* code; `jb_last_error()` gives
 * the message (mapped to `SimulatorError` by the Python host, like the reference's
 * june/activity/activity_manager.py:400-402 and june/simulator.py:588-607).
 */
#ifndef JUNE_B200_H
#define JUNE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JB_ABI_VERSION 1

/* ---- error codes ---------------------------------------------------------------------- */
#define JB_OK 0
#define JB_EINVAL 1   /* bad argument / inconsistent world description            */
#define JB_ECUDA 2    /* CUDA runtime failure (message in jb_last_error)          */
#define JB_ENOACT 3   /* "Some people do not have an activity in this timestep"   */
#define JB_ECAP 4     /* a device buffer capacity was exceeded                    */
#define JB_ENOGPU 5   /* no usable CUDA device: the product path has NO fallback  */

/* ---- activities, in the reference's hierarchy order (activity_manager.py:24-33) ------- */
#define JB_ACT_MEDICAL 0   /* medical_facility */
#define JB_ACT_COMMUTE 1   /* commute          */
#define JB_ACT_PRIMARY 2   /* primary_activity */
#define JB_ACT_LEISURE 3   /* leisure          */
#define JB_ACT_RESIDENCE 4 /* residence        */
#define JB_ACT_NONE 7      /* dead / nowhere   */
#define JB_N_ACT 5

/* ---- per-person state byte (`pstate`) -------------------------------------------------- */
#define JB_PS_ACT_MASK 0x07u /* activity chosen in the current step                        */
#define JB_PS_INFECTED 0x08u /* person.infection is not None (person.py:155-157)           */
#define JB_PS_DEAD 0x10u     /* person.dead                                                */
#define JB_PS_SEVERE 0x20u   /* infection.tag in severe_symptoms_stay_at_home_stage        */
#define JB_PS_MEDICAL 0x40u  /* person.subgroups.medical_facility is not None              */
#define JB_PS_SPARE 0x80u

/* ---- registered-member info word (`reg_info`, uint16) ---------------------------------- */
/* bits 0-7: local subgroup index inside the group; bits 8-10: activity that feeds the slot */
#define JB_INFO_LSG(x) ((x) & 0xFFu)
#define JB_INFO_ACT(x) (((x) >> 8) & 0x7u)
#define JB_MAKE_INFO(lsg, act) ((uint16_t)(((act) << 8) | (lsg)))

/* ---- group info word (`grp_info`, uint32) ---------------------------------------------- */
/* bits 0-7 beta index (row of JbStepParams.beta), 8-15 beta-scale index (company sector),
 * 16-31 region index */
#define JB_GINFO_BETA(x) ((x) & 0xFFu)
#define JB_GINFO_SCALE(x) (((x) >> 8) & 0xFFu)
#define JB_GINFO_REGION(x) ((x) >> 16)

/* ---- supergroup kinds (select the kernel specialisation) ------------------------------- */
#define JB_KIND_MATRIX 0 /* fixed contact matrix of dim<=JB_MAX_DIM (interactive.py:179-180) */
#define JB_KIND_SCHOOL 1 /* per-school matrix from 32x32 table + years[] (school.py:462-485) */

#define JB_MAX_DIM 8        /* largest fixed contact matrix (university 5x5)               */
#define JB_SCHOOL_DIM 32    /* school.py:436-438: teachers + ages 0..30                    */
#define JB_MAX_LSG 255      /* subgroups per group                                         */
#define JB_MAX_STAGES 8     /* longest covid19.yaml trajectory has 7 stages                */
#define JB_MAX_SUPERGROUPS 32
#define JB_MAX_VARIANTS 4
#define JB_N_TPAR 5         /* transmission parameters per infection                       */
#define JB_DYN_BIN_CAP 4096 /* dynamic members (patients) per group and step                */

/* transmission profile types (infection_selector.py:130-143) */
#define JB_TRANS_CONSTANT 0 /* par0 = probability                                          */
#define JB_TRANS_GAMMA 1    /* par = shape, shift, scale, norm (transmission.py:109-160)   */
#define JB_TRANS_XNEXP 2    /* par = n, alpha, time_first_infectious, norm, norm_time      */

/* completion-time distribution types (trajectory_maker.py:43-58) */
#define JB_DIST_CONSTANT 0    /* p0 = value                      */
#define JB_DIST_EXPONENTIAL 1 /* p0 = loc, p1 = scale            */
#define JB_DIST_BETA 2        /* p0 = a, p1 = b, p2 = loc, p3 = scale */
#define JB_DIST_LOGNORMAL 3   /* p0 = s, p1 = loc, p2 = scale    */
#define JB_DIST_NORMAL 4      /* p0 = loc, p1 = scale            */
#define JB_DIST_EXPONWEIB 5   /* p0 = a, p1 = c, p2 = loc, p3 = scale */

typedef struct {
  int32_t type;
  double p[4];
} JbDist;

/* One supergroup = one contiguous range of groups of one spec, processed in this order
 * (reference: simulator.py:399-455 iterates supergroups in activity-hierarchy order). */
typedef struct {
  uint32_t first_group, n_groups;
  int32_t kind;          /* JB_KIND_*                                                      */
  int32_t dim;           /* contact matrix dimension                                       */
  int32_t cm_offset;     /* offset (in doubles) of the dim*dim row-major matrix in `cm`    */
  int32_t activity;      /* activity whose presence in the step activates this supergroup  */
  int32_t launch_mode;   /* 0 = one thread per group, 1 = one warp per group               */
  int32_t has_dynamic;   /* groups of this supergroup may receive dynamic members: 0 no,
                          * 1 yes with JB_DYN_BIN_CAP entries per group and step, > 1 = that
                          * capacity (rounded up to a power of two)                         */
} JbSupergroup;

/* Disease description: symptom trajectories + transmission sampler + outcome table.
 * (covid19.yaml; trajectory_maker.py:215-227; infection_selector.py:280-332;
 *  health_index.py:151-182) */
typedef struct {
  int32_t n_tags;                 /* number of outcome slots (max tag value + 1), <= 16     */
  int32_t tag_recovered_mask;     /* bit (tag+2) set if tag is a recovered_stage            */
  int32_t tag_dead_mask;          /* bit (tag+2) set if tag is a fatality_stage             */
  int32_t tag_hospital_mask;      /* hospitalised_stage                                     */
  int32_t tag_icu_mask;           /* intensive_care_stage                                   */
  int32_t tag_severe_mask;        /* severe_symptoms_stay_at_home_stage                     */
  int32_t tag_exposed;            /* default_lowest_stage value                             */
  int32_t n_traj;                 /* number of trajectories                                 */
  int32_t traj_max_tag[16];       /* most severe tag of trajectory k (its key)              */
  int32_t traj_n_stages[16];
  int32_t traj_stage_tag[16][JB_MAX_STAGES];
  JbDist traj_stage_time[16][JB_MAX_STAGES];
  int32_t trans_type;             /* JB_TRANS_*                                             */
  JbDist trans_par[6];            /* gamma: max_infectiousness, shape, rate, shift          */
                                  /* xnexp: max_probability, smearing_time_first_infectious,
                                            smearing_peak_position, alpha, norm_time       */
                                  /* constant: probability                                  */
  /* cumulative outcome probabilities cum[pop(2: gp,ch)][sex(2: f,m)][age(100)][n_tags]     */
  const double* health_index_cum;
  uint32_t infection_id;          /* adler32(class name), infection.py:42-47                */
} JbDisease;

/* Everything needed to create a world.  Index spaces: persons [0,n_people) are residents of
 * this domain; [n_people, n_people+n_halo) are "halo" persons owned by another domain that are
 * registered members of groups here (the device form of mpi_wrapper.py:169-190 visitors). */
typedef struct {
  uint32_t abi_version;
  uint32_t n_people, n_halo, n_groups, n_members, n_supergroups, n_regions, n_super_areas;
  uint32_t n_variants, n_beta, n_scale, n_cm_doubles, n_school_years;
  uint32_t slot_capacity;          /* infection slots; 0 -> n_people                        */
  uint32_t newinf_capacity;        /* new infections per step; 0 -> n_people + n_halo       */
  uint64_t person_id_base;         /* global id of local person 0                           */
  /* persons (length n_people) */
  const uint8_t* age;              /* 0..99                                                 */
  const uint8_t* sex;              /* 0 = f, 1 = m                                          */
  const uint8_t* care_home_resident; /* 1 if residence is a care_home (health_index.py:160) */
  const uint8_t* has_activity;     /* bit a set: person has a static subgroup for activity a */
  const uint32_t* super_area;      /* person's home super area                              */
  /* super areas */
  const int32_t* sa_hospital;      /* closest_hospitals[0] as a group index, or -1          */
  const uint16_t* sa_region;
  /* groups, laid out supergroup after supergroup */
  const uint32_t* grp_offset;      /* n_groups + 1, CSR into reg_member                     */
  const uint32_t* grp_info;        /* JB_GINFO_*                                            */
  const uint32_t* grp_aux;         /* school: (years offset << 8) | n_years ; else 0        */
  const uint32_t* reg_member;      /* n_members person indices (local or halo)              */
  const uint16_t* reg_info;        /* n_members JB_MAKE_INFO(lsg, act)                      */
  const uint8_t* school_years;     /* n_school_years                                        */
  const JbSupergroup* supergroups;
  const double* cm;                /* all processed raw contact matrices                    */
  const double* beta_scale;        /* n_scale multipliers (company.py:309-313); [0] = 1.0   */
  JbDisease disease;
} JbWorldDesc;

/* Per-step inputs (the h2d payload of every step). */
typedef struct {
  double now;                 /* timer.now, days                                            */
  double delta_time;          /* timer.duration, days                                       */
  uint32_t step;              /* step ordinal (Philox counter word)                         */
  uint32_t activity_mask;     /* bit a set if activity a is in timer.activities             */
  uint64_t seed;              /* Philox key                                                 */
  uint32_t flags;             /* JB_STEP_*                                                  */
  uint32_t leisure_table;     /* 0 = no leisure probabilities this step, t + 1 = table t of
                               * jb_set_leisure_tables                                      */
  /* processed beta per (beta index, region): get_processed_beta, interactive.py:154-177,
   * household.py:256-309, school.py:487-522.  Row-major [n_beta][n_regions]. */
  const double* beta;
} JbStepParams;

#define JB_STEP_SEVERE_STAY_HOME 0x1u /* SevereSymptomsStayHome active (individual_policies.py:121-149) */
#define JB_STEP_HOSPITALISATION 0x2u  /* Hospitalisation active (medical_care_policies.py:413-526)     */
#define JB_STEP_INJECT_UNIFORMS 0x4u  /* Bernoulli draws come from jb_inject_uniforms               */
#define JB_STEP_NO_NEW_INFECTIONS 0x8u /* do not create infections on device (host uploads rows)    */
#define JB_STEP_RECORD_LAMBDA 0x10u   /* keep per-person Lambda of this step for jb_fetch_lambda   */
#define JB_STEP_INJECT_LEISURE 0x20u  /* leisure venues come from jb_inject_leisure (test mode)    */
#define JB_STEP_INJECT_POLICY 0x40u   /* policy uniforms come from jb_inject_policy_uniforms        */
#define JB_STEP_RECORD_EVENTS 0x80u   /* keep {where, infector} of every new infection (records)    */
#define JB_STEP_INJECT_COMMUTE 0x100u /* transports come from jb_inject_commute (test mode)         */
#define JB_STEP_RECORD_HEALTH 0x200u  /* keep the health events only (jb_fetch_health_events): what a caller needs to
                                       * mirror deaths / recoveries / symptom changes on its own objects, without the
                                       * cost of recording who infected whom (JB_STEP_RECORD_EVENTS implies it)  */

/* Per-step outputs (the d2h payload of every step). */
typedef struct {
  uint32_t n_new_infections;
  uint32_t n_infected;        /* currently infected after the health update                 */
  uint32_t n_dead_total;
  uint32_t n_recovered_step;
  uint32_t n_deaths_step;
  uint32_t n_hospitalised;    /* currently in ward + icu                                    */
  uint32_t n_icu;
  uint32_t n_draws;           /* Bernoulli draws consumed (== injected uniforms consumed)   */
  uint32_t n_no_activity;     /* >0 -> JB_ENOACT                                            */
  uint32_t n_slots_used;
  uint32_t n_halo_infected;   /* new infections of halo persons (to be sent home)           */
  uint32_t n_by_activity[JB_N_ACT]; /* residents placed per activity this step               */
} JbStepStats;

/* One host-created infection (jb_upload_infections), mirroring the fields the reference's
 * checkpoint stores (hdf5_savers/infection_savers/{transmission,symptoms}_saver.py). */
typedef struct {
  uint32_t person;            /* local person index                                         */
  uint32_t variant;           /* variant index                                              */
  double start_time;
  double tpar[JB_N_TPAR];
  double probability;         /* transmission.probability at upload time                    */
  int32_t n_stages;
  int32_t stage;
  int32_t max_tag;
  int32_t tag;
  double traj_time[JB_MAX_STAGES];
  int32_t traj_tag[JB_MAX_STAGES];
} JbInfectionRow;

typedef struct JbWorld JbWorld;

const char* jb_last_error(void);
int jb_abi_version(void);
int jb_device_count(void);

/* Create a device-resident world on CUDA device `device`.  Copies everything it needs. */
int jb_world_create(const JbWorldDesc* desc, int device, JbWorld** out);
void jb_world_destroy(JbWorld* w);

/* Replace person state wholesale (n_people entries each; NULL = keep).  `susceptibility` is
 * [n_variants][n_people].  Used for initial immunity (immunity_setter) and checkpoint restore. */
int jb_set_person_state(JbWorld* w, const uint8_t* pstate, const double* susceptibility);

/* Test mode: uniforms consumed by the Bernoulli draws in the reference's traversal order
 * (interaction.py:554-569,:636: one `random()` per susceptible of every group that must
 * timestep).  Copied to the device; consumed from index 0 by the next jb_step. */
int jb_inject_uniforms(JbWorld* w, const double* u, size_t n);

/* Infect people now using the device-side sampler (InfectionSelector._make_infection,
 * infection_selector.py:146-204): seeding (infection_seed.py:209-278). */
int jb_infect(JbWorld* w, const uint32_t* persons, const uint32_t* variants, size_t n,
              double time, uint64_t seed, uint32_t step);

/* Activities (bit a = JB_ACT_a) under which halo persons are registered in ANY domain; every rank must pass
 * the same value (OR over the ranks of the activities of registrations whose member is a halo person).  A step
 * whose activity mask does not intersect it exchanges nothing: commuters who are at home are present in no
 * group of another domain (the reference still sends its empty MovablePeople buffers,
 * activity_manager.py:404-447), and such a step may be run with jb_step on any world.  Default: all activities. */
int jb_set_halo_activities(JbWorld* w, uint32_t activity_mask);
/* Variants (infection classes): per-variant transmission-parameter distributions trans_par[n_variants][6]
 * (the reference gives every infection class its own selector, infection_selector.py:44-144; b117.yaml differs
 * from covid19.yaml in max_infectiousness) and cross immunity (Infection.immunity_ids, infection.py:136-191):
 * bit u of cross_immunity[v] set = an infection with variant v leaves the person immune to variant u.
 * Either pointer may be null.  Defaults: every variant samples from JbDisease.trans_par and protects against all. */
int jb_set_variants(JbWorld* w, uint32_t n_variants, const JbDist* trans_par, const uint32_t* cross_immunity);
/* Mutation event (event/mutation.py:9-66): every current infection of a resident of region r becomes `variant`
 * with probability regional_probability[r] (uniform: Philox stream 7 of (seed, step, person)): a complete new
 * infection of that variant is sampled and only its transmission is kept -- symptoms and start time stay, the
 * transmission probability restarts at its initial value until the next health update. */
int jb_mutate(JbWorld* w, uint32_t variant, const double* regional_probability, uint64_t seed, uint32_t step);
/* Restore person.subgroups.medical_facility: medical[n_people] = (hospital group << 8 | ward subgroup) or -1.
 * The reference's checkpoint does not carry it (patients are simply re-admitted by the next
 * Hospitalisation.apply, hdf5_savers/checkpoint_saver.py:162-218); it is restorable here so that a resumed
 * run continues bit for bit.  Call before jb_upload_infections. */
int jb_set_medical(JbWorld* w, const int32_t* medical);
/* Install host-created infections (the reference's own sampler, or a checkpoint). */
int jb_upload_infections(JbWorld* w, const JbInfectionRow* rows, size_t n);

/* One do_timestep: placement -> halo pack (if n_halo) ... -> interaction -> new infections ->
 * health update.  Split in two halves so that the host can run the NCCL halo exchange
 * between them (single-GPU callers use jb_step).
 *   jb_step_begin : placement + pack outbound halo state into d_send (may be NULL)
 *   jb_step_end   : unpack inbound halo state from d_recv, interaction, infection, health;
 *                   writes infected-abroad bytes into d_ret_send (may be NULL)
 *   jb_step_finish: applies d_ret_recv (infections of our residents abroad), health update.
 */
int jb_step(JbWorld* w, const JbStepParams* params, JbStepStats* stats_out);
/* Set-up aid: build whatever jb_step would build on the FIRST step with these activities / flags (the CUDA
 * library captures and instantiates the step's graph) without running anything, so that no later jb_step
 * pays for it.  The reference has no counterpart (its step has no set-up cost); a no-op is a valid
 * implementation. */
int jb_step_prepare(JbWorld* w, const JbStepParams* params);
int jb_step_begin(JbWorld* w, const JbStepParams* params, void* d_send);
int jb_step_interact(JbWorld* w, const void* d_recv, void* d_ret_send);
int jb_step_finish(JbWorld* w, const void* d_ret_recv, JbStepStats* stats_out);

/* Halo wiring (multi-GPU).  `out_person[i]` is the local person whose state is sent in
 * outbound slot i; slots are grouped by destination rank by the caller.  Inbound slot j
 * fills halo person n_people + j.  Record sizes in bytes are returned by jb_halo_record_size. */
int jb_set_halo(JbWorld* w, const uint32_t* out_person, size_t n_out);
size_t jb_halo_record_size(void);   /* bytes per person in d_send / d_recv       */
size_t jb_halo_return_size(void);   /* bytes per person in d_ret_send / d_ret_recv */

/* ---- leisure (groups/leisure/leisure.py:205-275, social_venue_distributor.py:257-278) ----
 * K leisure activities in the reference's distributor order.  Activity k sends a person to one
 * of the venues its AREA lists for k (`area.social_venues[spec]`), as a dynamic member of local
 * subgroup `subgroup[k]` of a group of supergroup `supergroup[k]` (which must have has_dynamic
 * set).  `area_off` is a CSR over (area, activity), index area * K + k, into `area_venue` (group
 * indices).  Care-home residents never do leisure (leisure.py:246-248). */
#define JB_MAX_LEISURE 8
typedef struct {
  uint32_t n_activities;       /* K <= JB_MAX_LEISURE                                        */
  uint32_t n_areas;
  const int32_t* supergroup;   /* [K]                                                        */
  const int32_t* subgroup;     /* [K] leisure_subgroup_type                                  */
  const uint32_t* person_area; /* [n_people]                                                 */
  const uint32_t* area_off;    /* [n_areas * K + 1]                                          */
  const uint32_t* area_venue;  /* [area_off[n_areas * K]] group indices                      */
} JbLeisureDesc;
int jb_set_leisure(JbWorld* w, const JbLeisureDesc* desc);

/* ---- residence visits (groups/leisure/residence_visits.py:297-345, groups/household.py:110-149,202-205) ----
 * One of the K leisure activities (`activity`, set up by jb_set_leisure with any supergroup: it is not used) sends a
 * person to a residence its OWN household is linked to: a household (the visitor joins the subgroup of its age --
 * kids < 18, young adults <= 25, adults < 65, old adults -- household.py:110-123) or a care home (subgroup
 * `care_home_subgroup` of a group of `care_home_supergroup`, which must have has_dynamic set).  Which kind, when the
 * household has both: residence_type_probabilities, normalised over the kinds it has (:308-322); which one: uniform.
 * A visited household welcomes its visitors (Household.get_leisure_subgroup -> make_household_residents_stay_home):
 * residents already out doing leisure come back home, residents not yet placed stay home instead of drawing a leisure
 * activity -- the reference walks the people in world order, and so does this (by the caller's person index).
 * CSR rows are indexed by the household's ordinal inside `household_supergroup` (group - first group). */
typedef struct {
  int32_t activity;              /* index of the residence_visits activity in the leisure tables          */
  int32_t household_supergroup;
  int32_t care_home_supergroup;  /* -1: no care-home visits                                                */
  int32_t care_home_subgroup;
  const uint32_t* household_off; /* [n_households + 1]                                                     */
  const uint32_t* household_target; /* group indices of `household_supergroup`                            */
  const uint32_t* care_home_off; /* [n_households + 1] or null                                             */
  const uint32_t* care_home_target;
  double p_household, p_care_home;  /* residence_type_probabilities                                        */
  int32_t household_visits_beta; /* beta rule of a household that is being visited ("household_visits",
                                  * InteractiveHousehold.get_processed_beta, household.py:279-289)        */
} JbVisitsDesc;
int jb_set_visits(JbWorld* w, const JbVisitsDesc* desc);
/* ChangeVisitsProbability (policy/leisure_policies.py:180-220): new residence_type_probabilities. */
int jb_set_visit_probabilities(JbWorld* w, double p_household, double p_care_home);
/* Probability tables (Leisure.generate_leisure_probabilities_for_timestep, leisure.py:205-228,
 * 297-372): table t holds does[n_regions][2 (f, m)][100] = P(does any activity) and
 * cum[n_regions][2][100][K] = cumulative P(activity | does).  JbStepParams.leisure_table picks
 * one per step. */
int jb_set_leisure_tables(JbWorld* w, uint32_t n_tables, const double* does, const double* cum);
/* Test mode (JB_STEP_INJECT_LEISURE): the venue of every person for the next step, -1 = none;
 * replaces the three random choices of leisure.py:253-263 / social_venue_distributor.py:257-266. */
int jb_inject_leisure(JbWorld* w, const int32_t* group_per_person);

/* ---- commute (groups/travel/travel.py:125-131, geography/city.py:83-102, station.py:52-76) ----
 * A public-transport commuter either commutes inside its work city (a random station of the city,
 * then a random transport of that station) or through the inter-city station of its home super
 * area (a random transport of that station); always subgroup 0 of the transport.  Transports are
 * groups of supergroups with has_dynamic set and activity JB_ACT_COMMUTE.
 * person_commute[p]: -1 = does not commute by public transport; (city << 1) = internal commuter
 * of that city; (station << 1) | 1 = commuter of that inter-city station. */
typedef struct {
  uint32_t n_cities, n_stations;
  const int32_t* person_commute;           /* [n_people]                                     */
  const uint32_t* city_station_off;        /* [n_cities + 1] CSR into city_station           */
  const uint32_t* city_station;            /* station indices (city.city_stations)           */
  const uint32_t* station_transport_off;   /* [n_stations + 1] CSR into station_transport    */
  const uint32_t* station_transport;       /* group indices (station.*_transports)           */
} JbCommuteDesc;
int jb_set_commute(JbWorld* w, const JbCommuteDesc* desc);
/* Test mode (JB_STEP_INJECT_COMMUTE): the transport of every person for the next step, -1 = none;
 * replaces the randint calls of city.py:95-97 and station.py:52-53,73-76. */
int jb_inject_commute(JbWorld* w, const int32_t* group_per_person);

/* ---- vaccination (epidemiology/vaccines/vaccination_campaign.py:36-385, vaccines.py:58-617;
 * called once per simulated day per living person from update_health_status,
 * epidemiology.py:295-303).  All dates are whole days (every day starts at the same hour), counted
 * from the simulation's first day. */
#define JB_MAX_VACCINES 8
#define JB_MAX_CAMPAIGNS 16
#define JB_MAX_DOSES 4
typedef struct {
  int32_t n_doses;                                   /* doses the vaccine defines (<= JB_MAX_DOSES) */
  int32_t days_administered_to_effective[JB_MAX_DOSES];
  int32_t days_effective_to_waning[JB_MAX_DOSES];
  int32_t days_waning[JB_MAX_DOSES];
  double waning_factor;
  const double* sterilisation; /* [n_doses][n_variants][100 ages] efficacy against infection   */
  const double* symptomatic;   /* [n_doses][n_variants][100 ages] efficacy against symptoms    */
} JbVaccine;
#define JB_VG_AGE 0        /* lo <= age < hi                         (vaccination_campaign.py:173-179) */
#define JB_VG_RESIDENCE 1  /* lo: 1 = care_home, 0 = household        (residence.group.spec)           */
#define JB_VG_PRIMARY 2    /* lo = JB_PA_KIND code of primary_activity.group.spec                      */
#define JB_VG_SEX 3        /* lo: 0 = f, 1 = m                                                         */
typedef struct {
  int32_t vaccine;                       /* index into the vaccine table                              */
  int32_t group_by, lo, hi;              /* JB_VG_*                                                   */
  int32_t n_doses;                       /* len(dose_numbers)                                         */
  int32_t dose_numbers[JB_MAX_DOSES];
  int32_t days_to_next_dose[JB_MAX_DOSES];
  int32_t start_day, end_day;            /* campaign active while start_day <= today < end_day        */
  uint32_t last_dose_vaccines;           /* bit v: previous vaccine v accepted; 0 = any (:197-199)    */
  double coverage;                       /* group_coverage                                            */
} JbVaccinationCampaign;
int jb_set_vaccination(JbWorld* w, const JbVaccine* vaccines, uint32_t n_vaccines,
                       const JbVaccinationCampaign* campaigns, uint32_t n_campaigns);
/* VaccinationCampaigns.apply + VaccineTrajectory.update_vaccine_effect for every living person,
 * for the simulated day `today` (enqueued on the world's stream, after the last step). */
int jb_vaccinate(JbWorld* w, int32_t today, uint64_t seed);
/* Test mode: the uniform each person consumes in VaccinationCampaigns.apply (:373) on the next
 * jb_vaccinate call; NULL switches back to Philox. */
int jb_inject_vaccination_uniforms(JbWorld* w, const double* u);
/* dose: person.vaccinated (-1 = None); vaccine: index of person.vaccine_type (-1 = None);
 * effective_multiplier: [n_variants][n_people] (immunity.effective_multiplier_dict). */
int jb_fetch_vaccination(JbWorld* w, int8_t* dose, int8_t* vaccine, double* effective_multiplier);

/* The whole per-person vaccination state, for checkpoints (the reference pickles person.vaccinated, vaccine_type and
 * vaccine_trajectory with the person, hdf5_savers/population_saver.py:261-278): read it (store = 0) or write it back
 * (store = 1, after jb_set_vaccination).  All arrays in the caller's person order; the three double arrays are
 * [n_variants][n_people].  Before jb_set_vaccination a read returns the "nobody vaccinated" state. */
typedef struct JbVaccinationState {
  int8_t* dose;             /* person.vaccinated (-1 = None)                                   */
  int8_t* vaccine;          /* index of person.vaccine_type (-1 = None)                        */
  int8_t* campaign;         /* campaign of the running vaccine_trajectory (-1 = None)          */
  int32_t* first_day;       /* day the trajectory started                                      */
  uint8_t* stage;           /* dose of the trajectory reached so far                           */
  double* prior_infection;  /* Dose.prior_efficacy of the running dose (vaccines.py:109-150)   */
  double* prior_symptoms;
  double* effective_multiplier;
} JbVaccinationState;
int jb_vaccination_state(JbWorld* w, int store, const JbVaccinationState* s);

/* ---- individual policies (policy/individual_policies.py:36-84) ---------------------------
 * The active individual policies, in the reference's list order, evaluated per person by the
 * placement stage.  A stay-home policy that fires reduces the person's activities to
 * (medical_facility, residence) and ends the evaluation (:52-77); a skip-activity policy removes
 * activities (:78-80).  Probabilities that are None in the reference are passed as NaN. */
#define JB_MAX_POLICIES 8
#define JB_POLICY_DRAWS 3          /* uniforms a policy may consume per person and step      */
#define JB_POL_SEVERE_STAY_HOME 1  /* SevereSymptomsStayHome (:121-149)                      */
#define JB_POL_QUARANTINE 2        /* Quarantine (:152-237): p0 n_days, p1 compliance,
                                    * mask = stay_at_home tags (bit tag+2)                    */
#define JB_POL_SHIELDING 3         /* Shielding (:416-439): p0 min_age, p1 compliance | NaN   */
#define JB_POL_CLOSE_SCHOOLS 4     /* CloseSchools (:480-530): flags bit0 full_closure,
                                    * mask = ages (0..31) of years_to_close, p0 attending_compliance */
#define JB_POL_CLOSE_UNIVERSITIES 5 /* CloseUniversities (:533-548)                          */
#define JB_POL_CLOSE_COMPANIES 6   /* CloseCompanies (:604-770): flags bit0 full_closure, p0
                                    * avoid_work_probability, p1 furlough_probability, p2
                                    * key_probability, p3..p5 furlough / key / random ratios    */
#define JB_POL_CLOSE_COMPANIES_TIERS 8 /* CloseCompaniesLockdownTiers (:551-601): mask bit r = region r is in lockdown
                                        * tier 3 or 4; a company worker with lockdown_status "random" whose work super
                                        * area is not the home one (jb_set_person_work_regions) skips work with
                                        * probability regional_compliance[home] when the work or the home region is
                                        * in such a tier */
#define JB_POL_LIMIT_LONG_COMMUTE 7 /* LimitLongCommute (:773-824): p0 going_to_work_probability; applies
                                    * to persons flagged JB_PA_LONG_COMMUTER.  As in the reference,
                                    * the activity is skipped when random() < p0.               */
typedef struct {
  int32_t type;
  uint32_t flags;
  uint32_t mask;
  uint32_t _pad;
  double p[6];
} JbIndividualPolicy;
/* work_region[n_people]: region of person.work_super_area when it differs from the home super area, else 0xFF
 * (CloseCompaniesLockdownTiers, individual_policies.py:570-599). */
int jb_set_person_work_regions(JbWorld* w, const uint8_t* work_region);
/* regional_compliance: [n_regions] (region.regional_compliance, regional_compliance.py:21-30) */
int jb_set_individual_policies(JbWorld* w, const JbIndividualPolicy* policies, uint32_t n,
                               const double* regional_compliance);
/* Static per-person facts the policies read: bits 0-1 lockdown_status (0 none, 1 key_worker,
 * 2 random, 3 furlough; person.py), bits 2-4 spec of the primary-activity group (0 none, 1 school,
 * 2 company, 3 university, 4 other), bit 5 = at least two key workers among the household's
 * residents (CloseSchools._check_kid_goes_to_school, :497-510), bit 6 = long-distance commuter. */
#define JB_PA_LOCKDOWN(x) ((x) & 3u)
#define JB_PA_KIND(x) (((x) >> 2) & 7u)
#define JB_PA_TWO_KEY_WORKERS 0x20u
#define JB_PA_LONG_COMMUTER 0x40u  /* person.id in LimitLongCommute.long_distance_commuter_ids (:795-813) */
int jb_set_person_attributes(JbWorld* w, const uint8_t* attr);
/* Test mode (JB_STEP_INJECT_POLICY): the uniforms the policies consume, [JB_MAX_POLICIES]
 * [JB_POLICY_DRAWS][n_people]; replaces Python's random() in individual_policies.py. */
int jb_inject_policy_uniforms(JbWorld* w, const double* u);

/* ---- halo exchange over peer memory (NVLink P2P) instead of host-issued collectives ----------
 * One process per GPU.  Every rank allocates an exchange window (inbound person records, inbound
 * "infected abroad" bytes, arrival flags), exports it as a CUDA IPC handle and opens its peers'
 * windows.  From then on jb_step runs halo worlds as ONE replayed graph: the pack kernel stores
 * the outbound records straight into the destination GPU's window and raises that GPU's flag;
 * the consumer spins on its own flag (bounded) before unpacking.  The two exchanges of a step
 * fence each other, so one window per direction suffices.
 *   jb_p2p_window : allocate the window of this world; writes 80 bytes: its 64-byte IPC handle
 *                   followed by two uint64 section sizes
 *   jb_p2p_connect: `handles` = the 64-byte handles of all ranks, [world_size][64], followed by the
 *                   section sizes of all ranks, [world_size][2] uint64; out_counts / in_counts
 *                   as for the all-to-all; peer_recv_base[q] = first record of this rank's segment
 *                   in rank q's inbound window, peer_ret_base[q] = first byte of this rank's
 *                   segment in rank q's return window (= q's outbound offset for this rank). */
#define JB_MAX_RANKS 16
int jb_p2p_window(JbWorld* w, void* ipc_handle_out);
int jb_p2p_connect(JbWorld* w, int rank, int world_size, const void* handles, const uint32_t* out_counts,
                   const uint32_t* in_counts, const uint32_t* peer_recv_base, const uint32_t* peer_ret_base);

int jb_sync(JbWorld* w);
/* Diagnostic: text summary of the device layout (kernel flavour per supergroup) into buf[cap]. */
int jb_world_describe(JbWorld* w, char* buf, size_t cap);

/* Readback (synchronises).  `cap` = capacity of the caller's arrays, `*n` = entries written. */
int jb_fetch_new_infections(JbWorld* w, uint32_t* member, uint32_t* variant, size_t cap, size_t* n);
int jb_fetch_person_state(JbWorld* w, uint8_t* pstate, int8_t* tag, double* trans_prob,
                          double* susceptibility, int32_t* medical);
int jb_fetch_lambda(JbWorld* w, double* lambda_out); /* n_people + n_halo, NaN where no draw */
int jb_fetch_infections(JbWorld* w, JbInfectionRow* rows, size_t cap, size_t* n);
/* Infection events of the last step run with JB_STEP_RECORD_EVENTS: the rows of the reference's
 * "infections" record table (interaction.py:664-683, records/event_records_writer.py:60-84) --
 * who was infected, in which group, with which variant, and the infector chosen by
 * _blame_subgroup / _blame_individuals (interaction.py:643-662).  Person indices are local
 * (halo persons: n_people + halo index); sorted by infected person. */
int jb_fetch_infection_events(JbWorld* w, uint32_t* infected, uint32_t* infector, uint32_t* group,
                              uint32_t* variant, size_t cap, size_t* n);
/* Health events of the last step run with JB_STEP_RECORD_EVENTS: rows of the reference's `deaths`
 * (epidemiology.py:176-186: location = hospital if the person had a medical facility, else the
 * residence), `recoveries` (:209-214), `hospital_admissions` (status ward_admitted only),
 * `icu_admissions` (status icu_transferred only) and `discharges` tables
 * (medical_care_policies.py:463-515, hospital.py:62-122).  group = -1 where the table has no location.
 * Sorted by (person, kind). */
#define JB_HEV_DEATH 1
#define JB_HEV_RECOVERY 2
#define JB_HEV_HOSPITAL_ADMISSION 3
#define JB_HEV_ICU_ADMISSION 4
#define JB_HEV_DISCHARGE 5
#define JB_HEV_SYMPTOMS 6 /* `symptoms` table (epidemiology.py:256-266): the tag changed; group = new tag */
int jb_fetch_health_events(JbWorld* w, uint32_t* person, uint32_t* kind, int32_t* group, size_t cap, size_t* n);
/* per-region summary of the last step: counts[n_regions][JB_N_SUMMARY], the quantities of a summary.csv row
 * (records_writer.py:151-164,248-376,764-773) with the reference's attribution:
 *   0 infections created in this domain, by the person's region (the reference's daily_infected goes by the region of
 *     the group an infection happened in: the host derives it from the infection events, Simulator.summary_rows);
 *   1 currently infected, by the person's region;
 *   2 / 3 patients placed in a ward / in intensive care by this step's placement, by the HOSPITAL's region
 *     (len(hospital.ward), len(hospital.icu): people admitted during the step's health update count from the next step);
 *   4 deaths, by the person's region;   5 deaths in hospital, by the hospital's region;   6 recoveries (person's region);
 *   7 ward admissions ("ward_admitted") and 8 transfers to intensive care ("icu_transferred"), by the hospital's region
 *     -- the two events the reference records (medical_care_policies.py:463-475). */
#define JB_N_SUMMARY 9
int jb_fetch_summary(JbWorld* w, uint32_t* counts);

/* Kernel timing hooks for bench.py: CUDA-event time (ms) of the interaction kernels and of
 * the whole step, measured on the library's stream, accumulated since the last reset. */
/* on: 0 off; 1 whole-step events only (CUDA-graph replay stays on); 2 additionally per-kernel
 * events around the interaction kernels (plain launches, no graph). */
int jb_timing_enable(JbWorld* w, int on);
/* on = 1 (default): the interaction kernels of different supergroups run concurrently on side
 * streams; on = 0: serially on the main stream (isolated per-kernel timings). */
int jb_set_concurrency(JbWorld* w, int on);
int jb_timing_read(JbWorld* w, double* interaction_ms, double* step_ms, uint64_t* n_steps,
                   uint64_t* n_launches, int reset);
/* Benchmark aid: overwrite `bytes` of scratch on the library's stream (evicts the L2 between timed
 * steps; enqueued in front of, i.e. outside, the next step's event bracket; no host wait). */
int jb_flush_l2(JbWorld* w, size_t bytes);
/* Algorithmic bytes of the interaction kernels for the last step. */
int jb_last_step_bytes(JbWorld* w, uint64_t* interaction_bytes, uint64_t* step_bytes);

#ifdef __cplusplus
}
#endif
#endif /* JUNE_B200_H */
