// libjune_b200.so -- C ABI (include/june_b200.h) over the kernels in jb_kernels.cuh.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cub/cub.cuh>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <limits>

#include "jb_kernels.cuh"
#include "jb_layout.cuh"

static thread_local std::string g_err;
void jb_set_error(const std::string& msg) { g_err = msg; }

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      jb_set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                       \
      return JB_ECUDA;                                                                        \
    }                                                                                         \
  } while (0)

#define REQUIRE(cond, msg)                                                                    \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      jb_set_error(msg);                                                                      \
      return JB_EINVAL;                                                                       \
    }                                                                                         \
  } while (0)

// JB_DEBUG_SYNC=1: synchronise after every kernel launch and name the kernel that faulted.
static bool jb_debug_sync() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("JB_DEBUG_SYNC"); v = (e && *e && *e != '0') ? 1 : 0; }
  return v == 1;
}
#define KCHECK(name, st)                                                                      \
  do {                                                                                        \
    CK(cudaGetLastError());                                                                   \
    if (jb_debug_sync()) {                                                                    \
      cudaError_t e_ = cudaStreamSynchronize(st);                                             \
      if (e_ != cudaSuccess) {                                                                \
        jb_set_error(std::string("kernel ") + (name) + ": " + cudaGetErrorString(e_));          \
        return JB_ECUDA;                                                                      \
      }                                                                                       \
    }                                                                                         \
  } while (0)

static int jb_env_flag(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

template <typename T>
static int dev_alloc(T** p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  CK(cudaMalloc((void**)p, n * sizeof(T)));
  return JB_OK;
}
template <typename T>
static int dev_upload(T** p, const T* src, size_t n, cudaStream_t s) {
  int rc = dev_alloc(p, n);
  if (rc) return rc;
  if (n && src) CK(cudaMemcpyAsync(*p, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
  return JB_OK;
}
#define TRY(x)            \
  do {                    \
    int rc_ = (x);        \
    if (rc_) return rc_;  \
  } while (0)

extern "C" const char* jb_last_error(void) { return g_err.c_str(); }
extern "C" int jb_abi_version(void) { return JB_ABI_VERSION; }
extern "C" int jb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
extern "C" size_t jb_halo_record_size(void) { return 0; }  // depends on V: see jb_halo_record_size_w
extern "C" size_t jb_halo_return_size(void) { return 1; }
extern "C" size_t jb_halo_record_size_w(JbWorld* w) { return 8 + (size_t)8 * w->V; }

static void copy_disease(const JbDisease& d, DiseaseDev& o) {
  memset(&o, 0, sizeof(o));
  o.n_tags = d.n_tags;
  o.rec_mask = d.tag_recovered_mask;
  o.dead_mask = d.tag_dead_mask;
  o.hosp_mask = d.tag_hospital_mask;
  o.icu_mask = d.tag_icu_mask;
  o.severe_mask = d.tag_severe_mask;
  o.tag_exposed = d.tag_exposed;
  o.n_traj = d.n_traj;
  for (int k = 0; k < 16; ++k) {
    o.traj_max_tag[k] = d.traj_max_tag[k];
    o.traj_n_stages[k] = d.traj_n_stages[k];
    for (int s = 0; s < JB_MAX_STAGES; ++s) {
      o.traj_stage_tag[k][s] = d.traj_stage_tag[k][s];
      o.traj_stage_time[k][s] = d.traj_stage_time[k][s];
    }
  }
  o.trans_type = d.trans_type;
  for (int k = 0; k < 6; ++k) o.trans_par[k] = d.trans_par[k];
  for (int v = 0; v < JB_MAX_VARIANTS; ++v) {
    for (int k = 0; k < 6; ++k) o.trans_par_v[v][k] = d.trans_par[k];
    o.cross[v] = 0xFFFFFFFFu;
  }
}

extern "C" int jb_world_create(const JbWorldDesc* d, int device, JbWorld** out) {
  if (!d || !out) { jb_set_error("null argument"); return JB_EINVAL; }
  *out = nullptr;
  REQUIRE(d->abi_version == JB_ABI_VERSION, "ABI version mismatch");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    jb_set_error("no CUDA device: june_b200 has no CPU fallback");
    return JB_ENOGPU;
  }
  REQUIRE(device >= 0 && device < ndev, "bad device index");
  REQUIRE(d->n_variants >= 1 && d->n_variants <= JB_MAX_VARIANTS, "n_variants out of range");
  REQUIRE(d->n_supergroups >= 1 && d->n_supergroups <= JB_MAX_SUPERGROUPS, "n_supergroups out of range");
  REQUIRE(d->n_people + (uint64_t)d->n_halo < (1ull << 28), "n_people + n_halo must be < 2^28");
  REQUIRE(d->n_groups < (1u << 27), "n_groups must be < 2^27 per domain");
  REQUIRE(d->disease.n_tags >= 1 && d->disease.n_tags <= 16 && d->disease.n_traj >= 1 && d->disease.n_traj <= 16,
          "disease table sizes out of range");
  CK(cudaSetDevice(device));
  JbWorld* w = new JbWorld();
  w->device = device;
  CK(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
  cudaStream_t s = w->stream;
  w->N_ext = d->n_people; w->H = d->n_halo; w->G = d->n_groups; w->M = d->n_members;
  w->V = d->n_variants; w->R = std::max(1u, d->n_regions); w->SA = d->n_super_areas;
  w->n_beta = d->n_beta; w->n_scale = d->n_scale; w->id_base = d->person_id_base;
  w->sgs.assign(d->supergroups, d->supergroups + d->n_supergroups);
  copy_disease(d->disease, w->dis_host);

  // validate supergroups + compute the max local-subgroup count (shared-memory sizing)
  std::vector<uint8_t> years(d->school_years, d->school_years + d->n_school_years);
  uint32_t expect = 0;
  for (size_t k = 0; k < w->sgs.size(); ++k) {
    const JbSupergroup& sg = w->sgs[k];
    REQUIRE(sg.first_group == expect, "supergroups must tile the group range contiguously");
    expect += sg.n_groups;
    int L = sg.dim;
    if (sg.kind == JB_KIND_SCHOOL) {
      REQUIRE(sg.dim == JB_SCHOOL_DIM, "school supergroup needs the 32x32 table");
      L = 1;
      for (uint32_t g = sg.first_group; g < sg.first_group + sg.n_groups; ++g) {
        uint32_t aux = d->grp_aux[g];
        L = std::max<int>(L, (int)(aux & 0xFF) + 1);
        REQUIRE((aux >> 8) + (aux & 0xFF) <= d->n_school_years, "school years out of range");
      }
    } else {
      REQUIRE(sg.dim >= 1 && sg.dim <= JB_MAX_DIM, "contact matrix dimension out of range");
      if (sg.launch_mode == 0) REQUIRE(sg.dim <= 4 && !sg.has_dynamic, "tile mode needs dim<=4 and static members");
    }
    REQUIRE(sg.cm_offset >= 0 && (uint32_t)(sg.cm_offset + sg.dim * sg.dim) <= d->n_cm_doubles, "cm offset out of range");
    REQUIRE(sg.activity >= 0 && sg.activity < JB_N_ACT, "bad supergroup activity");
    w->sg_L.push_back(L);
    {
      uint32_t acts = 0;
      for (uint32_t m = d->grp_offset[sg.first_group]; m < d->grp_offset[sg.first_group + sg.n_groups]; ++m)
        acts |= 1u << JB_INFO_ACT(d->reg_info[m]);
      w->sg_static_acts.push_back(acts);
    }
    {
      uint8_t hh = 0;
      for (uint32_t m = d->grp_offset[sg.first_group]; m < d->grp_offset[sg.first_group + sg.n_groups] && !hh; ++m)
        if (d->reg_member[m] >= d->n_people) hh = 1;
      w->sg_has_halo.push_back(hh);
    }
    if (sg.has_dynamic) w->any_dynamic = true;
  }
  REQUIRE(expect == w->G, "supergroups do not cover all groups");

  // internal numbering + tiles (jb_layout.cuh)
  HostLayout lay;
  {
    const char* err = jb_build_layout(d, w->sgs, (uint32_t)std::min(32, std::max(1, jb_env_flag("JB_TILE_MAX_GATHER", 16))),
                                      jb_env_flag("JB_MIRROR", 1) != 0, jb_env_flag("JB_SPAN", 1) != 0,
                                      (uint32_t)std::max(512, jb_env_flag("JB_SPAN_BIG_MAX", 4096)), lay);
    REQUIRE(err == nullptr, err ? err : "");
  }
  w->N = lay.N_int; w->NP = w->N + w->H;
  REQUIRE((uint64_t)w->NP < (1ull << 28), "internal person count must be < 2^28");
  w->C = d->slot_capacity ? d->slot_capacity : std::max(1u, w->N_ext);
  w->newinf_cap = d->newinf_capacity ? d->newinf_capacity : std::max(1u, w->N_ext + w->H);
  w->h_ext2int = lay.ext2int; w->h_int2ext = lay.int2ext;
  TRY(dev_upload(&w->int2ext, lay.int2ext.data(), w->N, s));
  TRY(dev_upload(&w->ext2int, lay.ext2int.data(), w->N_ext, s));
  w->tiles.resize(w->sgs.size());
  for (size_t k = 0; k < w->sgs.size(); ++k) {
    const HostTiles& T = lay.tiles[k];
    SgTiles& D = w->tiles[k];
    D.mode = T.mode;
    // groups the tiles do not cover, longest first (the long ones set the tail of the launch)
    const JbSupergroup& sg = w->sgs[k];
    std::vector<uint32_t> rest;
    if (T.mode == 0) { rest.resize(sg.n_groups); for (uint32_t i = 0; i < sg.n_groups; ++i) rest[i] = sg.first_group + i; }
    else rest = T.big;
    auto gsize = [&](uint32_t g) { return d->grp_offset[g + 1] - d->grp_offset[g]; };
    std::stable_sort(rest.begin(), rest.end(), [&](uint32_t x, uint32_t y) { return gsize(x) > gsize(y); });
    std::vector<uint32_t> lw, lc;
    // JB_WARP_MAX: largest group one warp processes (the warp keeps <= 128 members in registers and streams the rest)
    static const uint32_t warp_max = (uint32_t)std::max(32, jb_env_flag("JB_WARP_MAX", 4096));
    // A mirrored supergroup with one subgroup per group (1x1 contact matrix: companies) is processed by k_span, whose
    // work items the layout has built (jb_pack_span); only groups above JB_SPAN_BIG_MAX slots are left to the CTA list.
    const bool span = T.span;
    static const uint32_t span_big_max = (uint32_t)std::max(512, jb_env_flag("JB_SPAN_BIG_MAX", 4096));
    for (uint32_t g : rest) {
      if (span) { if (gsize(g) > span_big_max) lc.push_back(g); }
      else ((sg.has_dynamic || gsize(g) > warp_max) ? lc : lw).push_back(g);
    }
    D.n_w = (uint32_t)lw.size(); D.n_c = (uint32_t)lc.size();
    if (D.n_w) TRY(dev_upload(&D.list_w, lw.data(), lw.size(), s));
    if (D.n_c) TRY(dev_upload(&D.list_c, lc.data(), lc.size(), s));
    std::vector<uint32_t> lws, lcs;
    if (T.mode == 3) {  // mirrored: first slot of every listed group
      std::vector<uint32_t> slot_of(sg.n_groups, 0u);
      for (size_t i = 0; i < T.big.size(); ++i) slot_of[T.big[i] - sg.first_group] = T.big_slot0[i];
      for (uint32_t g : lw) lws.push_back(slot_of[g - sg.first_group]);
      for (uint32_t g : lc) lcs.push_back(slot_of[g - sg.first_group]);
      if (D.n_w) TRY(dev_upload(&D.list_w_slot, lws.data(), lws.size(), s));
      if (D.n_c) TRY(dev_upload(&D.list_c_slot, lcs.data(), lcs.size(), s));
      D.stream_base = lay.mir_base[k];
      if (span) {
        D.n_span = (uint32_t)T.span_items.size();
        if (D.n_span) TRY(dev_upload(&D.span_items, T.span_items.data(), T.span_items.size(), s));
        if (!T.span_group.empty()) TRY(dev_upload(&D.span_group, T.span_group.data(), T.span_group.size(), s));
      }
    }
    CK(cudaStreamSynchronize(s));  // lw / lc are locals
    if (T.mode == 0) continue;
    D.n_tiles = (uint32_t)(T.tp_info.size() / 32);
    REQUIRE(T.tile_group0.size() == D.n_tiles, "internal: tile table size mismatch");
    TRY(dev_upload(&D.tile_group0, T.tile_group0.data(), T.tile_group0.size(), s));
    {
      // the group word (beta rule | sector scale | region) of a tile whose groups all share it -- households of one
      // neighbourhood do: the tile kernel then reads it with the tile's other words, a stream, instead of gathering
      // one 32-byte sector per listed group
      std::vector<uint32_t> tgi(D.n_tiles, JB_TILE_GINFO_MIXED);
      for (uint32_t t = 0; t < D.n_tiles; ++t) {
        uint32_t g = T.tile_group0[t], word = JB_TILE_GINFO_MIXED;
        bool any = false, same = true;
        for (uint32_t x = 0; x < 32; ++x) {
          const uint8_t ti = T.tp_info[(size_t)t * 32 + x];
          if (!(ti & JB_TI_VALID) || !(ti & JB_TI_FIRST)) continue;
          const uint32_t wg = g < w->G ? d->grp_info[g] : JB_TILE_GINFO_MIXED;
          if (!any) { word = wg; any = true; } else if (wg != word) same = false;
          ++g;
        }
        if (any && same) tgi[t] = word;
      }
      TRY(dev_upload(&D.tile_ginfo, tgi.data(), tgi.size(), s));
      // what the descriptors say about a tile, as four words (k_tiles, LEAN): only for stream tiles whose slots are all
      // fed by one activity
      std::vector<uint4> tst;
      int uni = -1;
      bool uniform = T.mode == 2;
      for (size_t x = 0; uniform && x < T.tp_info.size(); ++x) {
        const uint8_t ti = T.tp_info[x];
        if (!(ti & JB_TI_VALID)) continue;
        if (uni < 0) uni = (int)JB_TI_ACT(ti); else if ((int)JB_TI_ACT(ti) != uni) uniform = false;
      }
      if (uniform && uni >= 0) {
        tst.resize(D.n_tiles);
        for (uint32_t t = 0; t < D.n_tiles; ++t) {
          uint4 m = make_uint4(0u, 0u, 0u, 0u);
          for (uint32_t x = 0; x < 32; ++x) {
            const uint8_t ti = T.tp_info[(size_t)t * 32 + x];
            if (ti & JB_TI_VALID) {
              m.x |= 1u << x;
              if (ti & JB_TI_FIRST) m.y |= 1u << x;
            }
            if (JB_TI_LSG(ti) & 1u) m.z |= 1u << x;
            if (JB_TI_LSG(ti) & 2u) m.w |= 1u << x;
          }
          tst[t] = m;
        }
        TRY(dev_upload(&D.tile_static, tst.data(), tst.size(), s));
        D.uniform_act = uni;
      }
      CK(cudaStreamSynchronize(s));  // tgi, tst are locals
    }
    TRY(dev_upload(&D.tp_info, T.tp_info.data(), T.tp_info.size(), s));
    if (T.mode == 1 || T.mode == 3) TRY(dev_upload(&D.tp_member, T.tp_member.data(), T.tp_member.size(), s));
    size_t n_first = 0, n_valid = 0;
    for (uint8_t x : T.tp_info) { n_first += (x & JB_TI_FIRST) ? 1 : 0; n_valid += (x & JB_TI_VALID) ? 1 : 0; }
    D.mean_len = n_first ? (double)n_valid / (double)n_first : 0.0;
    D.has_small = n_valid != 0;
  }
  if (lay.mir_slots) {
    // slot-ordered state mirror of the primary-activity supergroups.  `pmir` starts invalid, so the first
    // placement of a step with the primary activity pushes everybody's byte; halo slots are filled by the exchange.
    w->mir_slots = lay.mir_slots;
    TRY(dev_alloc(&w->mstate, (size_t)w->mir_slots + 16));
    TRY(dev_alloc(&w->mstep, (size_t)w->mir_slots + 16));
    TRY(dev_alloc(&w->pmir, (size_t)lay.N_int + 16));
    TRY(dev_upload(&w->pmslot, lay.pmslot.data(), lay.pmslot.size(), s));
    {
      // slot-ordered global ids (the Philox counter word of a draw): residents now, visitors in jb_set_halo_gids
      std::vector<uint32_t> mg((size_t)w->mir_slots, 0u);
      for (uint32_t p = 0; p < lay.N_int; ++p)
        if (lay.pmslot[p] != JB_MB_NONE) mg[lay.pmslot[p]] = (uint32_t)w->id_base + lay.int2ext[p];
      TRY(dev_upload(&w->mgid, mg.data(), mg.size(), s));
    }
    CK(cudaMemsetAsync(w->mstate, JB_MB_ABSENT, (size_t)w->mir_slots + 16, s));
    CK(cudaMemsetAsync(w->mstep, JB_MB_ABSENT, (size_t)w->mir_slots + 16, s));
    CK(cudaMemsetAsync(w->pmir, 0xFF, (size_t)lay.N_int + 16, s));
    CK(cudaStreamSynchronize(s));
  }
  TRY(dev_upload(&w->age, lay.age.data(), w->N, s));
  TRY(dev_upload(&w->sex, lay.sex.data(), w->N, s));
  TRY(dev_upload(&w->ch, lay.ch.data(), w->N, s));
  TRY(dev_upload(&w->phas, lay.phas.data(), w->N, s));
  TRY(dev_upload(&w->psa, lay.psa.data(), w->N, s));
  TRY(dev_upload(&w->sa_hospital, d->sa_hospital, w->SA, s));
  TRY(dev_upload(&w->sa_region, d->sa_region, w->SA, s));
  TRY(dev_upload(&w->grp_off, d->grp_offset, (size_t)w->G + 1, s));
  TRY(dev_upload(&w->grp_info, d->grp_info, w->G, s));
  TRY(dev_upload(&w->grp_aux, d->grp_aux, w->G, s));
  TRY(dev_upload(&w->reg_member, lay.reg_member.data(), w->M, s));
  TRY(dev_upload(&w->reg_info, d->reg_info, w->M, s));
  TRY(dev_upload(&w->school_years, d->school_years, d->n_school_years, s));
  TRY(dev_upload(&w->cm, d->cm, d->n_cm_doubles, s));
  TRY(dev_upload(&w->beta_scale, d->beta_scale, d->n_scale, s));
  TRY(dev_upload(&w->hi_cum, d->disease.health_index_cum, (size_t)400 * d->disease.n_tags, s));
  TRY(dev_upload(&w->dis, &w->dis_host, 1, s));

  TRY(dev_alloc(&w->pstate, w->NP));
  TRY(dev_alloc(&w->phot, w->NP + 16));
  TRY(dev_alloc(&w->pvar, w->NP));
  TRY(dev_alloc(&w->ptag, w->N));
  TRY(dev_alloc(&w->susc, (size_t)w->V * w->NP));
  TRY(dev_alloc(&w->trans[0], (size_t)2 * w->NP));   // [2][NP], one allocation (the half in use is a per-step scalar)
  w->trans[1] = w->trans[0] + w->NP;
  TRY(dev_alloc(&w->pmed, w->N));
  TRY(dev_alloc(&w->pslot, w->N));
  TRY(dev_alloc(&w->halo_gid, w->H));
  {
    // residents start alive, fully susceptible and nowhere; padding slots are permanently dead
    std::vector<uint8_t> st(w->NP, JB_SC_ONE);
    for (uint32_t i = 0; i < w->N; ++i)
      if (lay.int2ext[i] == 0xFFFFFFFFu) st[i] = JB_PS_DEAD;
    CK(cudaMemcpyAsync(w->pstate, st.data(), w->NP, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(w->phot, JB_ACT_NONE, w->NP + 16, s));
    CK(cudaStreamSynchronize(s));
  }
  CK(cudaMemsetAsync(w->pvar, 0, w->NP, s));
  CK(cudaMemsetAsync(w->ptag, 0xFF, w->N, s));  // -1 = healthy
  CK(cudaMemsetAsync(w->trans[0], 0, (size_t)w->NP * 8, s));
  CK(cudaMemsetAsync(w->trans[1], 0, (size_t)w->NP * 8, s));
  CK(cudaMemsetAsync(w->pmed, 0xFF, (size_t)w->N * 4, s));
  CK(cudaMemsetAsync(w->pslot, 0xFF, (size_t)w->N * 4, s));
  CK(cudaMemsetAsync(w->halo_gid, 0, (size_t)std::max(1u, w->H) * 4, s));
  {
    size_t n = (size_t)w->V * w->NP;
    k_fill_f64<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(w->susc, n, 1.0);
  }
  const size_t C = w->C;
  TRY(dev_alloc(&w->inf_person, C));
  TRY(dev_alloc(&w->inf_start, C));
  TRY(dev_alloc(&w->inf_par, C * JB_N_TPAR));
  TRY(dev_alloc(&w->inf_next, C));
  TRY(dev_alloc(&w->traj_time, C * JB_MAX_STAGES));
  TRY(dev_alloc(&w->inf_stage, C));
  TRY(dev_alloc(&w->inf_nst, C));
  TRY(dev_alloc(&w->inf_var, C));
  TRY(dev_alloc(&w->inf_tag, C));
  TRY(dev_alloc(&w->inf_maxtag, C));
  TRY(dev_alloc(&w->traj_tag, C * JB_MAX_STAGES));
  TRY(dev_alloc(&w->inf_region, C));
  TRY(dev_alloc(&w->inf_sev, C));
  TRY(dev_alloc(&w->inf_hosp, C));
  TRY(dev_alloc(&w->inf_med, C));
  CK(cudaMemsetAsync(w->inf_person, 0xFF, C * 4, s));

  // counters | per-region summary | dynamic-bin counts: one block, one memset per step
  {
    size_t slots = 0;
    for (size_t k = 0; k < w->sgs.size(); ++k) {
      if (!w->sgs[k].has_dynamic) continue;
      uint32_t cap = w->sgs[k].has_dynamic == 1 ? JB_DYN_BIN_CAP : (uint32_t)w->sgs[k].has_dynamic;
      uint32_t c2 = 32;
      while (c2 < cap) c2 <<= 1;  // the bin is sorted with a bitonic network
      REQUIRE(c2 <= 8192, "dynamic bin capacity above 8192 (shared-memory sort)");
      REQUIRE(w->dyn_ranges.n < JB_MAX_DYN_RANGES, "too many supergroups with dynamic members");
      REQUIRE(slots + (size_t)w->sgs[k].n_groups * c2 < (1ull << 32), "dynamic bins exceed 2^32 entries");
      w->sg_bin_base[k] = w->n_bins; w->sg_bin_slot[k] = (uint32_t)slots; w->sg_bin_cap[k] = c2;
      DynRanges& dr = w->dyn_ranges;
      const uint32_t r = dr.n++;
      dr.first[r] = w->sgs[k].first_group; dr.count[r] = w->sgs[k].n_groups;
      dr.cnt_base[r] = w->n_bins; dr.slot_base[r] = (uint32_t)slots; dr.cap[r] = c2;
      w->n_bins += w->sgs[k].n_groups;
      slots += (size_t)w->sgs[k].n_groups * c2;
    }
    if (slots) TRY(dev_alloc(&w->dyn_bin, slots));
  }
  w->counter_words = (size_t)CNT_N + (size_t)w->R * JB_N_SUMMARY + w->n_bins;
  TRY(dev_alloc(&w->counters, w->counter_words));
  CK(cudaMemsetAsync(w->counters, 0, w->counter_words * 4, s));
  w->summary = w->counters + CNT_N;
  w->dyn_cnt = w->summary + (size_t)w->R * JB_N_SUMMARY;
  {
    std::vector<uint8_t> reg(std::max(1u, w->N), 0);
    for (uint32_t p = 0; p < w->N_ext; ++p) reg[lay.ext2int[p]] = (uint8_t)d->sa_region[d->super_area[p]];
    TRY(dev_upload(&w->pregion, reg.data(), reg.size(), s));
    w->h_pregion = reg;
    w->h_phosp.assign(std::max(1u, w->N), -1);
    for (uint32_t p = 0; p < w->N_ext; ++p) w->h_phosp[lay.ext2int[p]] = d->sa_hospital[d->super_area[p]];
  }
  TRY(dev_alloc(&w->newinf_member, w->newinf_cap));
  TRY(dev_alloc(&w->newinf_var, w->newinf_cap));
  w->step_bytes = sizeof(StepDev) + sizeof(double) * (size_t)w->n_beta * w->R;
  TRY(dev_alloc(&w->d_step, w->step_bytes));
  TRY(dev_alloc(&w->d_step_aux, sizeof(StepDev)));
  w->beta = reinterpret_cast<double*>(w->d_step + sizeof(StepDev));
  TRY(dev_alloc(&w->draw_cnt, (size_t)w->G + 1));
  TRY(dev_alloc(&w->draw_base, (size_t)w->G + 1));
  {
    size_t b2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b2, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)(w->G + 1), s);
    w->cub_tmp_bytes = b2 + 256;
    CK(cudaMalloc(&w->cub_tmp, w->cub_tmp_bytes));
  }
  w->step_stride = (w->step_bytes + 63) & ~(size_t)63;
  CK(cudaHostAlloc((void**)&w->h_step_ring, 4 * w->step_stride, cudaHostAllocMapped));
  CK(cudaHostGetDevicePointer((void**)&w->d_step_ring, w->h_step_ring, 0));
  for (int i = 0; i < 4; ++i) {
    w->h_step[i] = w->h_step_ring + i * w->step_stride;
    CK(cudaEventCreateWithFlags(&w->ev_step_buf[i], cudaEventDisableTiming));
  }
  if (jb_env_flag("JB_STAMPS", 0)) {
    unsigned long long* st = nullptr;
    CK(cudaMalloc((void**)&st, JB_STAMP_RING * 12 * 8));
    CK(cudaMemset(st, 0xFF, JB_STAMP_RING * 8 * 8));
    CK(cudaMemset(st + JB_STAMP_RING * 8, 0, JB_STAMP_RING * 4 * 8));
    CK(cudaMemcpyToSymbol(g_stamps, &st, sizeof(st)));
    w->stamps = st;
  }
  TRY(dev_alloc(&w->d_step_seq, 1));
  CK(cudaMemsetAsync(w->d_step_seq, 0, 4, s));
  CK(cudaHostAlloc((void**)&w->h_counters, sizeof(uint32_t) * (CNT_N + (size_t)w->R * JB_N_SUMMARY), cudaHostAllocMapped));
  CK(cudaHostGetDevicePointer((void**)&w->d_h_counters, w->h_counters, 0));
  memset(w->h_counters, 0, sizeof(uint32_t) * (CNT_N + (size_t)w->R * JB_N_SUMMARY));
  CK(cudaEventCreate(&w->ev_step0));
  CK(cudaEventCreate(&w->ev_step1));
  CK(cudaEventCreate(&w->ev_int0));
  CK(cudaEventCreate(&w->ev_int1));
  for (size_t k = 0; k < w->sgs.size(); ++k) {
    CK(cudaEventCreate(&w->ev_sg0[k]));
    CK(cudaEventCreate(&w->ev_sg1[k]));
    for (int u = 0; u < 3; ++u) {
      CK(cudaStreamCreateWithFlags(&w->side[k][u], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&w->ev_join[k][u], cudaEventDisableTiming));
    }
  }
  CK(cudaEventCreateWithFlags(&w->ev_fork, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&w->side_health, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&w->ev_fork_health, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&w->ev_join_health, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&w->side_halo, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&w->ev_fork_halo, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&w->ev_halo_ready, cudaEventDisableTiming));
  // the CTA kernel may need > 48 KB of dynamic shared memory for very large schools
  CK(cudaFuncSetAttribute(k_groups<1, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_groups<1, JB_CTA_WARPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_groups<JB_MAX_VARIANTS, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_groups<JB_MAX_VARIANTS, JB_CTA_WARPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_groups<1, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_groups<1, JB_CTA_WARPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_groups<JB_MAX_VARIANTS, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_groups<JB_MAX_VARIANTS, JB_CTA_WARPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s));
  *out = w;
  return JB_OK;
}

extern "C" void jb_world_destroy(JbWorld* w) {
  if (!w) return;
  cudaSetDevice(w->device);
  cudaStreamSynchronize(w->stream);
  void* ptrs[] = {w->age, w->sex, w->ch, w->phas, w->school_years, w->psa, w->grp_off, w->grp_info, w->grp_aux,
                  w->reg_member, w->reg_info, w->sa_region, w->sa_hospital, w->cm, w->beta_scale, w->hi_cum, w->dis,
                  w->mstate, w->mstep, w->pmir, w->pmslot, w->mgid, w->pstate, w->phot, w->pvar, w->ptag, w->susc, w->trans[0], w->pmed, w->pslot, w->halo_gid, w->out_person,
                  w->inf_person, w->inf_start, w->inf_par, w->inf_next, w->traj_time, w->inf_stage, w->inf_nst,
                  w->inf_var, w->inf_tag, w->inf_maxtag, w->traj_tag, w->inf_region, w->inf_sev, w->inf_hosp, w->inf_med, w->counters, w->newinf_member, w->newinf_var,
                  w->d_step, w->d_step_aux, w->lambda, w->inject, w->draw_cnt, w->draw_base, w->dyn_bin, w->pregion,
                  w->cub_tmp, w->int2ext, w->ext2int, w->d_vacc, w->vac_dose, w->vac_type, w->traj_camp, w->traj_day0, w->traj_stage,
                  w->prior_eff_inf, w->prior_eff_sym, w->effmult, w->vac_inject, w->cm_person, w->cm_inject, w->cm_cs_off, w->cm_cs, w->cm_st_off, w->cm_st, w->ev_infected, w->ev_infector, w->ev_group, w->ev_var, w->hev_person, w->hev_kind, w->hev_group, w->pres_group, w->pol_subjects, w->pwork_region, w->d_pol, w->pattr, w->pol_inject, w->lz_parea, w->lz_off, w->lz_venue, w->lz_does, w->lz_cum, w->lz_inject};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (double* q : w->vacc_tables) cudaFree(q);
  for (void* q : {(void*)w->vz.hh_off, (void*)w->vz.hh, (void*)w->vz.ch_off, (void*)w->vz.ch, (void*)w->vz.choice, (void*)w->vz.list,
                  (void*)w->vz.hv[0], (void*)w->vz.vslot, (void*)w->vz.vis_group, (void*)w->vz.vis_cnt, (void*)w->vz.vis_bin})
    if (q) cudaFree(q);
  for (void* q : w->p2p_opened) cudaIpcCloseMemHandle(q);
  if (w->p2p_window) cudaFree(w->p2p_window);
  if (w->flush_buf) cudaFree(w->flush_buf);
  for (cudaEvent_t e : w->ring0) cudaEventDestroy(e);
  for (cudaEvent_t e : w->ring1) cudaEventDestroy(e);
  if (w->p2p_done) cudaFree(w->p2p_done);
  if (w->p2p_ret_local) cudaFree(w->p2p_ret_local);
  for (SgTiles& t : w->tiles) {
    if (t.tile_group0) cudaFree(t.tile_group0);
    if (t.tile_ginfo) cudaFree(t.tile_ginfo);
    if (t.tile_static) cudaFree(t.tile_static);
    if (t.tp_member) cudaFree(t.tp_member);
    if (t.tp_info) cudaFree(t.tp_info);
    if (t.list_w) cudaFree(t.list_w);
    if (t.list_c) cudaFree(t.list_c);
    if (t.list_w_slot) cudaFree(t.list_w_slot);
    if (t.list_c_slot) cudaFree(t.list_c_slot);
    if (t.span_items) cudaFree(t.span_items);
    if (t.span_group) cudaFree(t.span_group);
  }
  for (int i = 0; i < 4; ++i) {
    if (w->ev_step_buf[i]) cudaEventDestroy(w->ev_step_buf[i]);
  }
  for (JbWorld::GraphEntry& e : w->graphs) cudaGraphExecDestroy(e.exec);
  if (w->stamps) {
    // average start offsets (us, relative to k_step_begin) over the recorded steps, split by whether the step
    // ran a gather-tile / company kernel (primary-activity steps are longer)
    std::vector<unsigned long long> h(JB_STAMP_RING * 12);
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), w->stamps, h.size() * 8, cudaMemcpyDeviceToHost);
    double acc[8] = {0}, ext[4] = {0}; int n = 0; double total = 0; int nt = 0;
    for (int r = 0; r < JB_STAMP_RING; ++r) {
      const unsigned long long* q = &h[(size_t)r * 8];
      bool ok = q[0] != ~0ull;
      for (int k = 1; k < 8; ++k) ok = ok && q[k] != ~0ull && q[k] != 0ull;
      if (!ok) continue;
      for (int k = 0; k < 8; ++k) acc[k] += (double)(q[k] - q[0]) * 1e-3;
      for (int k = 0; k < 4; ++k) { const unsigned long long e = h[(size_t)JB_STAMP_RING * 8 + (size_t)r * 4 + k]; if (e > q[0]) ext[k] += (double)(e - q[0]) * 1e-3; }
      ++n;
      const unsigned long long* nx = &h[(size_t)((r + 1) % JB_STAMP_RING) * 8];
      if (nx[0] != ~0ull && nx[0] > q[0] && nx[0] - q[0] < 1000000ull) { total += (double)(nx[0] - q[0]) * 1e-3; ++nt; }
    }
    if (n) {
      fprintf(stderr, "[jb stamps] %d steps: place %.1f tiles %.1f groups_first %.1f groups_last %.1f health %.1f newinf %.1f end %.1f | begin-to-begin %.1f us (%d)\n",
              n, acc[1] / n, acc[2] / n, acc[3] / n, acc[4] / n, acc[5] / n, acc[6] / n, acc[7] / n, nt ? total / nt : 0.0, nt);
      fprintf(stderr, "[jb stamps] last block: tiles phase A done %.1f, tiles done %.1f, newinf done %.1f\n", ext[0] / n, ext[1] / n, ext[3] / n);
    }
    unsigned long long* null_ptr = nullptr;
    cudaMemcpyToSymbol(g_stamps, &null_ptr, sizeof(null_ptr));
    cudaFree(w->stamps);
  }
  if (w->h_counters) cudaFreeHost(w->h_counters);
  if (w->h_step_ring) cudaFreeHost(w->h_step_ring);
  if (w->d_step_seq) cudaFree(w->d_step_seq);
  if (w->ev_step0) cudaEventDestroy(w->ev_step0);
  if (w->ev_step1) cudaEventDestroy(w->ev_step1);
  if (w->ev_int0) cudaEventDestroy(w->ev_int0);
  if (w->ev_int1) cudaEventDestroy(w->ev_int1);
  for (int k = 0; k < JB_MAX_SUPERGROUPS; ++k) { if (w->ev_sg0[k]) cudaEventDestroy(w->ev_sg0[k]); if (w->ev_sg1[k]) cudaEventDestroy(w->ev_sg1[k]); }
  for (int k = 0; k < JB_MAX_SUPERGROUPS; ++k)
    for (int u = 0; u < 3; ++u) {
      if (w->side[k][u]) cudaStreamDestroy(w->side[k][u]);
      if (w->ev_join[k][u]) cudaEventDestroy(w->ev_join[k][u]);
    }
  if (w->ev_fork) cudaEventDestroy(w->ev_fork);
  if (w->side_health) cudaStreamDestroy(w->side_health);
  if (w->ev_fork_health) cudaEventDestroy(w->ev_fork_health);
  if (w->ev_join_health) cudaEventDestroy(w->ev_join_health);
  if (w->side_halo) cudaStreamDestroy(w->side_halo);
  if (w->ev_fork_halo) cudaEventDestroy(w->ev_fork_halo);
  if (w->ev_halo_ready) cudaEventDestroy(w->ev_halo_ready);
  cudaStreamDestroy(w->stream);
  delete w;
}

extern "C" int jb_set_person_state(JbWorld* w, const uint8_t* pstate, const double* susceptibility) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  if (pstate) {
    // ABI byte -> persistent flags (bits 3-6) and the activity of the (last) step
    std::vector<uint8_t> st(w->N, JB_PS_DEAD), hot(w->N, JB_ACT_NONE);
    for (uint32_t p = 0; p < w->N_ext; ++p) {
      st[w->h_ext2int[p]] = pstate[p] & 0x78u;
      hot[w->h_ext2int[p]] = (pstate[p] & (JB_PS_ACT_MASK | JB_PS_INFECTED));
    }
    CK(cudaMemcpy(w->pstate, st.data(), w->N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w->phot, hot.data(), w->N, cudaMemcpyHostToDevice));
    // the dead take part in nothing (Activities(None, ...)) and are counted
    std::vector<uint8_t> has(w->N);
    CK(cudaMemcpy(has.data(), w->phas, w->N, cudaMemcpyDeviceToHost));
    uint32_t n_dead = 0;
    for (uint32_t p = 0; p < w->N_ext; ++p)
      if (pstate[p] & JB_PS_DEAD) { has[w->h_ext2int[p]] = 0; ++n_dead; }
    CK(cudaMemcpy(w->phas, has.data(), w->N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w->counters + CNT_DEAD_TOTAL, &n_dead, 4, cudaMemcpyHostToDevice));
  }
  if (susceptibility) {
    std::vector<double> su(w->N, 1.0);
    for (uint32_t v = 0; v < w->V; ++v) {
      for (uint32_t p = 0; p < w->N_ext; ++p) su[w->h_ext2int[p]] = susceptibility[(size_t)v * w->N_ext + p];
      CK(cudaMemcpy(w->susc + (size_t)v * w->NP, su.data(), (size_t)w->N * 8, cudaMemcpyHostToDevice));
    }
  }
  if (w->N) {
    k_refresh_codes<<<(w->N + 255) / 256, 256, 0, w->stream>>>(w->pstate, w->susc, w->N);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(w->stream));
  }
  return JB_OK;
}

static void drop_graphs(JbWorld* w);
extern "C" int jb_set_variants(JbWorld* w, uint32_t n_variants, const JbDist* trans_par, const uint32_t* cross_immunity) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  REQUIRE(n_variants == w->V, "jb_set_variants: n_variants differs from the world's");
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  for (uint32_t v = 0; v < JB_MAX_VARIANTS; ++v) {
    for (int k = 0; k < 6; ++k)
      w->dis_host.trans_par_v[v][k] = (trans_par && v < n_variants) ? trans_par[(size_t)v * 6 + k] : w->dis_host.trans_par[k];
    w->dis_host.cross[v] = (cross_immunity && v < n_variants) ? cross_immunity[v] : 0xFFFFFFFFu;
  }
  CK(cudaMemcpy(w->dis, &w->dis_host, sizeof(w->dis_host), cudaMemcpyHostToDevice));
  drop_graphs(w);  // the disease tables travel as a kernel parameter
  return JB_OK;
}

extern "C" int jb_set_halo_activities(JbWorld* w, uint32_t activity_mask) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  w->halo_act_mask = activity_mask;
  drop_graphs(w);
  return JB_OK;
}

extern "C" int jb_set_medical(JbWorld* w, const int32_t* medical) {
  if (!w || !medical) { jb_set_error("null argument"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  std::vector<int32_t> med(w->N, -1);
  std::vector<uint8_t> st(w->N);
  CK(cudaMemcpy(st.data(), w->pstate, w->N, cudaMemcpyDeviceToHost));
  for (uint32_t p = 0; p < w->N_ext; ++p) {
    const uint32_t i = w->h_ext2int[p];
    med[i] = medical[p];
    st[i] = (uint8_t)(medical[p] >= 0 ? (st[i] | JB_PS_MEDICAL) : (st[i] & ~JB_PS_MEDICAL));
  }
  CK(cudaMemcpy(w->pmed, med.data(), (size_t)w->N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(w->pstate, st.data(), w->N, cudaMemcpyHostToDevice));
  return JB_OK;
}

extern "C" int jb_inject_uniforms(JbWorld* w, const double* u, size_t n) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  if (n > w->inject_cap) {
    CK(cudaStreamSynchronize(w->stream));
    if (w->inject) cudaFree(w->inject);
    w->inject_cap = n + n / 2 + 16;
    CK(cudaMalloc((void**)&w->inject, w->inject_cap * 8));
  }
  if (n) CK(cudaMemcpyAsync(w->inject, u, n * 8, cudaMemcpyHostToDevice, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  w->n_inject = n;
  return JB_OK;
}

// ---- leisure set-up -------------------------------------------------------------------------
static void drop_graphs(JbWorld* w) {  // captured launches hold kernel arguments by value
  for (JbWorld::GraphEntry& e : w->graphs) cudaGraphExecDestroy(e.exec);
  w->graphs.clear();
}

extern "C" int jb_set_leisure(JbWorld* w, const JbLeisureDesc* d) {
  if (!w || !d) { jb_set_error("null argument"); return JB_EINVAL; }
  REQUIRE(d->n_activities >= 1 && d->n_activities <= JB_MAX_LEISURE, "n_activities out of range");
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  drop_graphs(w);
  LeisureDev& z = w->lz;
  const uint32_t K = d->n_activities;
  for (uint32_t k = 0; k < K; ++k) {
    REQUIRE(d->supergroup[k] < 0 || ((size_t)d->supergroup[k] < w->sgs.size() && w->sgs[d->supergroup[k]].has_dynamic),
            "leisure supergroup must exist and have dynamic members");  // (-1: no venues -- residence visits, jb_set_visits)
    REQUIRE(d->subgroup[k] >= 0 && d->subgroup[k] < 16, "leisure subgroup out of range");
    z.sg_first[k] = d->supergroup[k] < 0 ? 0u : w->sgs[d->supergroup[k]].first_group;
    z.sg_count[k] = d->supergroup[k] < 0 ? 0u : w->sgs[d->supergroup[k]].n_groups;
    z.lsg[k] = d->subgroup[k];
  }
  const size_t n_off = (size_t)d->n_areas * K + 1;
  std::vector<uint32_t> pa(std::max(1u, w->N), 0);
  for (uint32_t p = 0; p < w->N_ext; ++p) {
    REQUIRE(d->person_area[p] < d->n_areas, "person_area out of range");
    pa[w->h_ext2int[p]] = d->person_area[p];
  }
  for (size_t i = 0; i < d->area_off[n_off - 1]; ++i) REQUIRE(d->area_venue[i] < w->G, "area_venue out of range");
  for (uint32_t* q : {w->lz_parea, w->lz_off, w->lz_venue}) if (q) cudaFree(q);
  TRY(dev_upload(&w->lz_parea, pa.data(), pa.size(), w->stream));
  TRY(dev_upload(&w->lz_off, d->area_off, n_off, w->stream));
  TRY(dev_upload(&w->lz_venue, d->area_venue, (size_t)d->area_off[n_off - 1], w->stream));
  CK(cudaStreamSynchronize(w->stream));
  z.K = K; z.parea = w->lz_parea; z.off = w->lz_off; z.venue = w->lz_venue;
  return JB_OK;
}

static int ensure_res_group(JbWorld* w) {
  if (w->pres_group) return JB_OK;
  TRY(dev_alloc(&w->pres_group, w->NP));
  CK(cudaMemsetAsync(w->pres_group, 0xFF, (size_t)w->NP * 4, w->stream));
  k_res_group<<<(w->G + 255) / 256, 256, 0, w->stream>>>(w->grp_off, w->reg_member, w->reg_info, w->G, w->N, w->pres_group);
  KCHECK("k_res_group", w->stream);
  return JB_OK;
}

extern "C" int jb_set_visits(JbWorld* w, const JbVisitsDesc* d) {
  if (!w || !d) { jb_set_error("null argument"); return JB_EINVAL; }
  REQUIRE(w->lz.K != 0, "jb_set_visits before jb_set_leisure");
  REQUIRE(d->activity >= 0 && (uint32_t)d->activity < w->lz.K && d->household_supergroup >= 0 &&
          (size_t)d->household_supergroup < w->sgs.size() && d->household_off && d->household_target, "jb_set_visits: bad argument");
  REQUIRE(d->care_home_supergroup < 0 || ((size_t)d->care_home_supergroup < w->sgs.size() && w->sgs[d->care_home_supergroup].has_dynamic),
          "care-home visits need a care-home supergroup with dynamic members");
  REQUIRE(d->household_visits_beta >= 0 && (uint32_t)d->household_visits_beta < w->n_beta, "household_visits_beta out of range");
  REQUIRE(d->p_household + d->p_care_home > 0.0, "residence_type_probabilities must not both be zero");
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  drop_graphs(w);
  const JbSupergroup& hs = w->sgs[d->household_supergroup];
  REQUIRE(w->sg_L[d->household_supergroup] >= 4, "the household supergroup needs the four age subgroups");
  VisitsDev& z = w->vz;
  for (void* q : {(void*)z.hh_off, (void*)z.hh, (void*)z.ch_off, (void*)z.ch, (void*)z.choice, (void*)z.list, (void*)z.hv[0], (void*)z.vslot,
                  (void*)z.vis_group, (void*)z.vis_cnt, (void*)z.vis_bin})
    if (q) cudaFree(q);
  z = VisitsDev();
  const uint32_t nh = hs.n_groups;
  for (uint32_t i = 0; i < d->household_off[nh]; ++i)
    REQUIRE(d->household_target[i] - hs.first_group < nh, "household_target outside the household supergroup");
  z.activity = d->activity; z.hh_first = hs.first_group; z.hh_count = nh;
  uint32_t *p1 = nullptr, *p2 = nullptr;
  TRY(dev_upload(&p1, d->household_off, (size_t)nh + 1, w->stream));
  TRY(dev_upload(&p2, d->household_target, std::max<size_t>(1, d->household_off[nh]), w->stream));
  z.hh_off = p1; z.hh = p2;
  if (d->care_home_supergroup >= 0 && d->care_home_off) {
    const JbSupergroup& cs = w->sgs[d->care_home_supergroup];
    for (uint32_t i = 0; i < d->care_home_off[nh]; ++i)
      REQUIRE(d->care_home_target[i] - cs.first_group < cs.n_groups, "care_home_target outside the care-home supergroup");
    z.ch_first = cs.first_group; z.ch_count = cs.n_groups; z.ch_lsg = (uint32_t)d->care_home_subgroup;
    uint32_t *p3 = nullptr, *p4 = nullptr;
    TRY(dev_upload(&p3, d->care_home_off, (size_t)nh + 1, w->stream));
    TRY(dev_upload(&p4, d->care_home_target, std::max<size_t>(1, d->care_home_off[nh]), w->stream));
    z.ch_off = p3; z.ch = p4;
  }
  z.p_household = d->p_household / (d->p_household + d->p_care_home);
  TRY(ensure_res_group(w));
  z.pres_group = w->pres_group;
  TRY(dev_alloc(&z.choice, std::max(1u, w->N)));
  TRY(dev_alloc(&z.list, std::max(1u, w->N)));
  TRY(dev_alloc(&z.hv[0], (size_t)3 * nh));
  z.hv[1] = z.hv[0] + nh; z.hv[2] = z.hv[1] + nh;
  TRY(dev_alloc(&z.vslot, nh));
  z.vis_cap = std::max(1u, nh);  // on a weekend afternoon a quarter of the households and more receive somebody
  TRY(dev_alloc(&z.vis_group, z.vis_cap));
  TRY(dev_alloc(&z.vis_cnt, z.vis_cap));
  TRY(dev_alloc(&z.vis_bin, (size_t)z.vis_cap * JB_VIS_BIN));
  CK(cudaMemsetAsync(z.choice, 0xFF, (size_t)std::max(1u, w->N) * 4, w->stream));
  CK(cudaMemsetAsync(z.hv[0], 0xFF, (size_t)3 * nh * 4, w->stream));
  CK(cudaMemsetAsync(z.vslot, 0xFF, (size_t)nh * 4, w->stream));
  CK(cudaMemsetAsync(z.vis_cnt, 0, (size_t)z.vis_cap * 4, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  w->visits_beta = d->household_visits_beta;
  w->visits_sg = d->household_supergroup;
  return JB_OK;
}

extern "C" int jb_set_visit_probabilities(JbWorld* w, double p_household, double p_care_home) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  REQUIRE(w->vz.hh_off != nullptr, "jb_set_visit_probabilities before jb_set_visits");
  REQUIRE(p_household + p_care_home > 0.0, "residence_type_probabilities must not both be zero");
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  drop_graphs(w);  // (the value is a kernel argument of the captured placement)
  w->vz.p_household = p_household / (p_household + p_care_home);
  return JB_OK;
}

extern "C" int jb_set_leisure_tables(JbWorld* w, uint32_t n_tables, const double* does, const double* cum) {
  if (!w || !does || !cum) { jb_set_error("null argument"); return JB_EINVAL; }
  REQUIRE(w->lz.K != 0, "jb_set_leisure_tables before jb_set_leisure");
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  drop_graphs(w);
  if (w->lz_does) cudaFree(w->lz_does);
  if (w->lz_cum) cudaFree(w->lz_cum);
  const size_t per = (size_t)w->R * 200;
  TRY(dev_upload(&w->lz_does, does, per * n_tables, w->stream));
  TRY(dev_upload(&w->lz_cum, cum, per * w->lz.K * n_tables, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  w->lz.does = w->lz_does; w->lz.cum = w->lz_cum; w->lz.n_tables = n_tables;
  return JB_OK;
}

extern "C" int jb_inject_leisure(JbWorld* w, const int32_t* group_per_person) {
  if (!w || !group_per_person) { jb_set_error("null argument"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  std::vector<int32_t> v(std::max(1u, w->N), -1);
  for (uint32_t p = 0; p < w->N_ext; ++p) v[w->h_ext2int[p]] = group_per_person[p];
  if (!w->lz_inject) { TRY(dev_alloc(&w->lz_inject, v.size())); drop_graphs(w); }
  CK(cudaMemcpy(w->lz_inject, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  return JB_OK;
}

template <typename T>
static int fetch_permuted(JbWorld* w, const T* dev, T* out);

// ---- peer-memory halo exchange ---------------------------------------------------------------------
// Window layout: [recv: H records][ret: n_out bytes, 256-aligned][flag_people: JB_MAX_RANKS u32][flag_ret: JB_MAX_RANKS u32]
static size_t p2p_align(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int jb_p2p_window(JbWorld* w, void* handle_out) {
  if (!w || !handle_out) { jb_set_error("null argument"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  if (!w->p2p_window) {
    w->p2p_recv_bytes = p2p_align((size_t)std::max(1u, w->H) * (8 + (size_t)8 * w->V));
    w->p2p_ret_bytes = p2p_align(std::max(1u, w->n_out));
    const size_t bytes = w->p2p_recv_bytes + w->p2p_ret_bytes + 2 * JB_MAX_RANKS * 4;
    CK(cudaMalloc((void**)&w->p2p_window, bytes));
    CK(cudaMemset(w->p2p_window, 0, bytes));
    TRY(dev_alloc(&w->p2p_done, 4));  // [0], [1] last-block tickets of the two push kernels, [2] sticky "peer lost"
    CK(cudaMemset(w->p2p_done, 0, 16));
    TRY(dev_alloc(&w->p2p_ret_local, std::max(1u, w->H)));
    CK(cudaMemset(w->p2p_ret_local, 0, std::max(1u, w->H)));
  }
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, w->p2p_window));
  static_assert(sizeof(h) == 64, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  // followed by the two section sizes of this rank's window (the peers need them to find the flags)
  const uint64_t sizes[2] = {w->p2p_recv_bytes, w->p2p_ret_bytes};
  memcpy((unsigned char*)handle_out + 64, sizes, 16);
  return JB_OK;
}

extern "C" int jb_p2p_connect(JbWorld* w, int rank, int ws, const void* handles, const uint32_t* out_counts,
                              const uint32_t* in_counts, const uint32_t* peer_recv_base, const uint32_t* peer_ret_base) {
  if (!w || !handles || !out_counts || !in_counts || !peer_recv_base || !peer_ret_base) { jb_set_error("null argument"); return JB_EINVAL; }
  REQUIRE(ws >= 1 && ws <= JB_MAX_RANKS && rank >= 0 && rank < ws, "bad rank / world size");
  REQUIRE(w->p2p_window != nullptr, "jb_p2p_connect before jb_p2p_window");
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  drop_graphs(w);
  P2PDev& P = w->p2p;
  memset(&P, 0, sizeof(P));
  P.rank = rank; P.ws = ws;
  uint32_t so = 0, si = 0;
  for (int q = 0; q < ws; ++q) {
    P.out_off[q] = so; P.in_off[q] = si;
    so += out_counts[q]; si += in_counts[q];
    P.peer_recv_base[q] = peer_recv_base[q]; P.peer_ret_base[q] = peer_ret_base[q];
  }
  for (int q = ws; q <= JB_MAX_RANKS; ++q) { P.out_off[q] = so; P.in_off[q] = si; }
  REQUIRE(so == w->n_out && si == w->H, "out_counts / in_counts do not match the halo of this world");
  for (int q = 0; q < ws; ++q) {
    unsigned char* base = w->p2p_window;
    if (q != rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const unsigned char*)handles + (size_t)q * 64, 64);
      void* ptr = nullptr;
      CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
      w->p2p_opened.push_back(ptr);
      base = (unsigned char*)ptr;
    }
    // every rank lays its window out with ITS OWN sizes: they travel in the first words of the handle blob
    const uint64_t* sizes = reinterpret_cast<const uint64_t*>((const unsigned char*)handles + (size_t)ws * 64 + (size_t)q * 16);
    P.recv[q] = base;
    P.ret[q] = base + sizes[0];
    P.flag_people[q] = reinterpret_cast<uint32_t*>(base + sizes[0] + sizes[1]);
    P.flag_ret[q] = P.flag_people[q] + JB_MAX_RANKS;
  }
  w->p2p_on = true;
  return JB_OK;
}

// ---- vaccination -------------------------------------------------------------------------------
extern "C" int jb_set_vaccination(JbWorld* w, const JbVaccine* vac, uint32_t nv, const JbVaccinationCampaign* cam, uint32_t nc) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  REQUIRE(nv <= JB_MAX_VACCINES && nc <= JB_MAX_CAMPAIGNS, "too many vaccines / campaigns");
  REQUIRE((nv == 0 || vac) && (nc == 0 || cam), "null argument");
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  drop_graphs(w);  // k_newinf starts reading the effective multipliers
  for (double* q : w->vacc_tables) cudaFree(q);
  w->vacc_tables.clear();
  VaccDev h;
  memset(&h, 0, sizeof(h));
  h.n_vaccines = nv; h.n_campaigns = nc;
  for (uint32_t k = 0; k < nv; ++k) {
    REQUIRE(vac[k].n_doses >= 1 && vac[k].n_doses <= JB_MAX_DOSES, "vaccine doses out of range");
    h.vaccines[k] = vac[k];
    const size_t n = (size_t)vac[k].n_doses * w->V * 100;
    double *d0 = nullptr, *d1 = nullptr;
    TRY(dev_upload(&d0, vac[k].sterilisation, n, w->stream));
    TRY(dev_upload(&d1, vac[k].symptomatic, n, w->stream));
    w->vacc_tables.push_back(d0); w->vacc_tables.push_back(d1);
    h.vaccines[k].sterilisation = d0; h.vaccines[k].symptomatic = d1;
  }
  for (uint32_t k = 0; k < nc; ++k) {
    const JbVaccinationCampaign& c = cam[k];
    REQUIRE(c.vaccine >= 0 && (uint32_t)c.vaccine < nv && c.n_doses >= 1 && c.n_doses <= JB_MAX_DOSES, "bad campaign");
    for (int i = 0; i < c.n_doses; ++i)
      REQUIRE(c.dose_numbers[i] >= 0 && c.dose_numbers[i] < vac[c.vaccine].n_doses, "campaign dose number not defined by its vaccine");
    h.campaigns[k] = c;
  }
  if (!w->d_vacc) TRY(dev_alloc(&w->d_vacc, 1));
  CK(cudaMemcpyAsync(w->d_vacc, &h, sizeof(h), cudaMemcpyHostToDevice, w->stream));
  if (!w->vac_dose) {
    const size_t N = std::max(1u, w->N);
    TRY(dev_alloc(&w->vac_dose, N)); TRY(dev_alloc(&w->vac_type, N)); TRY(dev_alloc(&w->traj_camp, N));
    TRY(dev_alloc(&w->traj_day0, N)); TRY(dev_alloc(&w->traj_stage, N));
    TRY(dev_alloc(&w->prior_eff_inf, N * w->V)); TRY(dev_alloc(&w->prior_eff_sym, N * w->V)); TRY(dev_alloc(&w->effmult, N * w->V));
    CK(cudaMemsetAsync(w->vac_dose, 0xFF, N, w->stream)); CK(cudaMemsetAsync(w->vac_type, 0xFF, N, w->stream));
    CK(cudaMemsetAsync(w->traj_camp, 0xFF, N, w->stream)); CK(cudaMemsetAsync(w->traj_day0, 0, N * 4, w->stream));
    CK(cudaMemsetAsync(w->traj_stage, 0, N, w->stream));
    CK(cudaMemsetAsync(w->prior_eff_inf, 0, N * w->V * 8, w->stream)); CK(cudaMemsetAsync(w->prior_eff_sym, 0, N * w->V * 8, w->stream));
    k_fill_f64<<<(unsigned)((N * w->V + 255) / 256), 256, 0, w->stream>>>(w->effmult, N * w->V, 1.0);
  }
  CK(cudaStreamSynchronize(w->stream));
  w->n_campaigns = nc;
  return JB_OK;
}

extern "C" int jb_vaccinate(JbWorld* w, int32_t today, uint64_t seed) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  REQUIRE(!w->in_step, "jb_vaccinate inside a step");
  if (!w->n_campaigns || !w->N) return JB_OK;
  CK(cudaSetDevice(w->device));
  VaccArgs a;
  memset(&a, 0, sizeof(a));
  a.tab = w->d_vacc; a.N = w->N; a.NP = w->NP; a.V = (int)w->V; a.today = today; a.seed = seed;
  a.gid_base = (uint32_t)w->id_base; a.int2ext = w->int2ext; a.pstate = w->pstate;
  a.age = w->age; a.sex = w->sex; a.ch = w->ch; a.pattr = w->pattr; a.susc = w->susc;
  a.vac_dose = w->vac_dose; a.vac_type = w->vac_type; a.traj_camp = w->traj_camp; a.traj_day0 = w->traj_day0;
  a.traj_stage = w->traj_stage; a.prior_eff_inf = w->prior_eff_inf; a.prior_eff_sym = w->prior_eff_sym; a.effmult = w->effmult;
  a.inject = w->vac_inject_on ? w->vac_inject : nullptr;
  k_vaccinate<<<(w->N + 255) / 256, 256, 0, w->stream>>>(a);
  KCHECK("k_vaccinate", w->stream);
  w->launches++;
  return JB_OK;
}

extern "C" int jb_inject_vaccination_uniforms(JbWorld* w, const double* u) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  w->vac_inject_on = u != nullptr;
  if (!u) return JB_OK;
  std::vector<double> v(std::max(1u, w->N), 2.0);
  for (uint32_t p = 0; p < w->N_ext; ++p) v[w->h_ext2int[p]] = u[p];
  if (!w->vac_inject) TRY(dev_alloc(&w->vac_inject, v.size()));
  CK(cudaMemcpy(w->vac_inject, v.data(), v.size() * 8, cudaMemcpyHostToDevice));
  return JB_OK;
}

extern "C" int jb_fetch_vaccination(JbWorld* w, int8_t* dose, int8_t* vaccine, double* effmult) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  if (!w->vac_dose) {
    for (uint32_t p = 0; p < w->N_ext; ++p) { if (dose) dose[p] = -1; if (vaccine) vaccine[p] = -1; }
    if (effmult) for (size_t i = 0; i < (size_t)w->N_ext * w->V; ++i) effmult[i] = 1.0;
    return JB_OK;
  }
  if (dose) TRY(fetch_permuted(w, w->vac_dose, dose));
  if (vaccine) TRY(fetch_permuted(w, w->vac_type, vaccine));
  if (effmult)
    for (uint32_t v = 0; v < w->V; ++v) TRY(fetch_permuted(w, w->effmult + (size_t)v * w->N, effmult + (size_t)v * w->N_ext));
  return JB_OK;
}

template <typename T>
static int store_permuted(JbWorld* w, T* dev, const T* in, T pad) {
  std::vector<T> tmp(w->N, pad);
  for (uint32_t p = 0; p < w->N_ext; ++p) tmp[w->h_ext2int[p]] = in[p];
  CK(cudaMemcpy(dev, tmp.data(), (size_t)w->N * sizeof(T), cudaMemcpyHostToDevice));
  return JB_OK;
}

extern "C" int jb_vaccination_state(JbWorld* w, int store, const JbVaccinationState* st) {
  if (!w || !st) { jb_set_error("null argument"); return JB_EINVAL; }
  REQUIRE(st->dose && st->vaccine && st->campaign && st->first_day && st->stage && st->prior_infection && st->prior_symptoms &&
          st->effective_multiplier, "jb_vaccination_state: null array");
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  if (!w->vac_dose) {
    REQUIRE(!store, "jb_vaccination_state: jb_set_vaccination first");
    for (uint32_t p = 0; p < w->N_ext; ++p) { st->dose[p] = -1; st->vaccine[p] = -1; st->campaign[p] = -1; st->first_day[p] = 0; st->stage[p] = 0; }
    for (size_t i = 0; i < (size_t)w->N_ext * w->V; ++i) { st->prior_infection[i] = 0.0; st->prior_symptoms[i] = 0.0; st->effective_multiplier[i] = 1.0; }
    return JB_OK;
  }
  if (!store) {
    TRY(fetch_permuted(w, w->vac_dose, st->dose)); TRY(fetch_permuted(w, w->vac_type, st->vaccine));
    TRY(fetch_permuted(w, w->traj_camp, st->campaign)); TRY(fetch_permuted(w, w->traj_day0, st->first_day));
    TRY(fetch_permuted(w, w->traj_stage, st->stage));
    for (uint32_t v = 0; v < w->V; ++v) {
      TRY(fetch_permuted(w, w->prior_eff_inf + (size_t)v * w->N, st->prior_infection + (size_t)v * w->N_ext));
      TRY(fetch_permuted(w, w->prior_eff_sym + (size_t)v * w->N, st->prior_symptoms + (size_t)v * w->N_ext));
      TRY(fetch_permuted(w, w->effmult + (size_t)v * w->N, st->effective_multiplier + (size_t)v * w->N_ext));
    }
    return JB_OK;
  }
  TRY(store_permuted<int8_t>(w, w->vac_dose, st->dose, -1)); TRY(store_permuted<int8_t>(w, w->vac_type, st->vaccine, -1));
  TRY(store_permuted<int8_t>(w, w->traj_camp, st->campaign, -1)); TRY(store_permuted<int32_t>(w, w->traj_day0, st->first_day, 0));
  TRY(store_permuted<uint8_t>(w, w->traj_stage, st->stage, 0));
  for (uint32_t v = 0; v < w->V; ++v) {
    TRY(store_permuted<double>(w, w->prior_eff_inf + (size_t)v * w->N, st->prior_infection + (size_t)v * w->N_ext, 0.0));
    TRY(store_permuted<double>(w, w->prior_eff_sym + (size_t)v * w->N, st->prior_symptoms + (size_t)v * w->N_ext, 0.0));
    TRY(store_permuted<double>(w, w->effmult + (size_t)v * w->N, st->effective_multiplier + (size_t)v * w->N_ext, 1.0));
  }
  return JB_OK;
}

// ---- commute ---------------------------------------------------------------------------------
extern "C" int jb_set_commute(JbWorld* w, const JbCommuteDesc* d) {
  if (!w || !d) { jb_set_error("null argument"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  drop_graphs(w);
  std::vector<int32_t> pc(std::max(1u, w->N), -1);
  for (uint32_t p = 0; p < w->N_ext; ++p) {
    const int32_t c = d->person_commute[p];
    REQUIRE(c < 0 || ((c & 1) ? (uint32_t)(c >> 1) < d->n_stations : (uint32_t)(c >> 1) < d->n_cities), "person_commute out of range");
    pc[w->h_ext2int[p]] = c;
  }
  for (uint32_t i = 0; i < d->city_station_off[d->n_cities]; ++i) REQUIRE(d->city_station[i] < d->n_stations, "city_station out of range");
  for (uint32_t i = 0; i < d->station_transport_off[d->n_stations]; ++i) REQUIRE(d->station_transport[i] < w->G, "station_transport out of range");
  for (void* q : {(void*)w->cm_person, (void*)w->cm_cs_off, (void*)w->cm_cs, (void*)w->cm_st_off, (void*)w->cm_st}) if (q) cudaFree(q);
  TRY(dev_upload(&w->cm_person, pc.data(), pc.size(), w->stream));
  TRY(dev_upload(&w->cm_cs_off, d->city_station_off, (size_t)d->n_cities + 1, w->stream));
  TRY(dev_upload(&w->cm_cs, d->city_station, (size_t)d->city_station_off[d->n_cities], w->stream));
  TRY(dev_upload(&w->cm_st_off, d->station_transport_off, (size_t)d->n_stations + 1, w->stream));
  TRY(dev_upload(&w->cm_st, d->station_transport, (size_t)d->station_transport_off[d->n_stations], w->stream));
  CK(cudaStreamSynchronize(w->stream));
  w->cmt.person = w->cm_person; w->cmt.cs_off = w->cm_cs_off; w->cmt.cs = w->cm_cs; w->cmt.st_off = w->cm_st_off; w->cmt.st = w->cm_st;
  return JB_OK;
}

extern "C" int jb_inject_commute(JbWorld* w, const int32_t* g) {
  if (!w || !g) { jb_set_error("null argument"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  std::vector<int32_t> v(std::max(1u, w->N), -1);
  for (uint32_t p = 0; p < w->N_ext; ++p) v[w->h_ext2int[p]] = g[p];
  if (!w->cm_inject) { TRY(dev_alloc(&w->cm_inject, v.size())); drop_graphs(w); }
  CK(cudaMemcpy(w->cm_inject, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  return JB_OK;
}

// ---- individual policies ---------------------------------------------------------------------
extern "C" int jb_set_individual_policies(JbWorld* w, const JbIndividualPolicy* pol, uint32_t n, const double* rc) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  REQUIRE(n <= JB_MAX_POLICIES, "too many individual policies");
  REQUIRE(n == 0 || w->R <= JB_MAX_REGIONS_POLICY, "too many regions for the policy table");
  REQUIRE(n == 0 || pol != nullptr, "null policies");
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  PolicyDev h;
  memset(&h, 0, sizeof(h));
  h.n = n;
  for (uint32_t k = 0; k < n; ++k) h.pol[k] = pol[k];
  for (uint32_t r = 0; r < w->R && r < JB_MAX_REGIONS_POLICY; ++r) h.rc[r] = rc ? rc[r] : 1.0;
  if (!w->d_pol) TRY(dev_alloc(&w->d_pol, 1));
  CK(cudaMemcpy(w->d_pol, &h, sizeof(h), cudaMemcpyHostToDevice));
  bool subject_mode = n > 0;
  for (uint32_t k = 0; k < n; ++k) subject_mode = subject_mode && pol[k].type == JB_POL_LIMIT_LONG_COMMUTE;
  if (w->n_pol != n || subject_mode != w->pol_subject_mode) drop_graphs(w);  // both shape the launch sequence
  w->n_pol = n;
  w->pol_subject_mode = subject_mode;
  return JB_OK;
}

extern "C" int jb_set_person_attributes(JbWorld* w, const uint8_t* attr) {
  if (!w || !attr) { jb_set_error("null argument"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  std::vector<uint8_t> v(std::max(1u, w->N), 0);
  for (uint32_t p = 0; p < w->N_ext; ++p) v[w->h_ext2int[p]] = attr[p];
  if (!w->pattr) TRY(dev_alloc(&w->pattr, v.size()));
  CK(cudaMemcpy(w->pattr, v.data(), v.size(), cudaMemcpyHostToDevice));
  // the static subject list of LimitLongCommute
  std::vector<uint32_t> subj;
  for (uint32_t i = 0; i < w->N; ++i)
    if (v[i] & JB_PA_LONG_COMMUTER) subj.push_back(i);
  if (w->pol_subjects) { cudaFree(w->pol_subjects); w->pol_subjects = nullptr; }
  w->n_pol_subjects = (uint32_t)subj.size();
  if (!subj.empty()) TRY(dev_upload(&w->pol_subjects, subj.data(), subj.size(), w->stream));
  CK(cudaStreamSynchronize(w->stream));
  drop_graphs(w);
  return JB_OK;
}

extern "C" int jb_set_person_work_regions(JbWorld* w, const uint8_t* work_region) {
  if (!w || !work_region) { jb_set_error("null argument"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  std::vector<uint8_t> v(std::max(1u, w->N), 0xFF);
  for (uint32_t p = 0; p < w->N_ext; ++p) v[w->h_ext2int[p]] = work_region[p];
  if (!w->pwork_region) { TRY(dev_alloc(&w->pwork_region, v.size())); drop_graphs(w); }
  CK(cudaMemcpy(w->pwork_region, v.data(), v.size(), cudaMemcpyHostToDevice));
  return JB_OK;
}

extern "C" int jb_inject_policy_uniforms(JbWorld* w, const double* u) {
  if (!w || !u) { jb_set_error("null argument"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  const size_t rows = (size_t)JB_MAX_POLICIES * JB_POLICY_DRAWS;
  std::vector<double> v(rows * std::max(1u, w->N), 2.0);
  for (size_t r = 0; r < rows; ++r)
    for (uint32_t p = 0; p < w->N_ext; ++p) v[r * w->N + w->h_ext2int[p]] = u[r * w->N_ext + p];
  if (!w->pol_inject) { TRY(dev_alloc(&w->pol_inject, v.size())); drop_graphs(w); }
  CK(cudaMemcpy(w->pol_inject, v.data(), v.size() * 8, cudaMemcpyHostToDevice));
  return JB_OK;
}

static inline const StepDev* step_dev(JbWorld* w) { return reinterpret_cast<const StepDev*>(w->d_step); }

static NewInfArgs make_newinf_args(JbWorld* w) {
  NewInfArgs a;
  memset(&a, 0, sizeof(a));
  a.N = w->N; a.NP = w->NP; a.C = w->C; a.V = (int)w->V; a.gid_base = (uint32_t)w->id_base; a.int2ext = w->int2ext;
  a.age = w->age; a.sex = w->sex; a.ch = w->ch; a.hi_cum = w->hi_cum; a.dis = w->dis;
  a.pstate = w->pstate; a.pvar = w->pvar; a.ptag = w->ptag; a.susc = w->susc; a.pslot = w->pslot;
  a.trans = w->trans[0];
  a.inf_person = w->inf_person; a.inf_start = w->inf_start; a.inf_par = w->inf_par; a.inf_next = w->inf_next;
  a.traj_time = w->traj_time; a.inf_stage = w->inf_stage; a.inf_nst = w->inf_nst; a.inf_var = w->inf_var;
  a.inf_tag = w->inf_tag; a.inf_maxtag = w->inf_maxtag; a.traj_tag = w->traj_tag; a.counters = w->counters;
  a.summary = w->summary; a.pregion = w->pregion;
  a.inf_region = w->inf_region; a.inf_hosp = w->inf_hosp; a.inf_med = w->inf_med; a.inf_sev = w->inf_sev;
  a.psa = w->psa; a.sa_hospital = w->sa_hospital; a.pmed = w->pmed;
  a.effmult = w->effmult;
  return a;
}

extern "C" int jb_infect(JbWorld* w, const uint32_t* persons, const uint32_t* variants, size_t n, double time,
                         uint64_t seed, uint32_t step) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  if (n == 0) return JB_OK;
  CK(cudaSetDevice(w->device));
  uint32_t* d_p = nullptr;
  uint8_t* d_v = nullptr;
  std::vector<uint8_t> v8(n, 0);
  std::vector<uint32_t> pint(n);
  for (size_t i = 0; i < n; ++i) {
    REQUIRE(persons[i] < w->N_ext, "jb_infect: person index out of range");
    pint[i] = w->h_ext2int[persons[i]];
    if (variants) { REQUIRE(variants[i] < w->V, "jb_infect: variant out of range"); v8[i] = (uint8_t)variants[i]; }
  }
  TRY(dev_upload(&d_p, pint.data(), n, w->stream));
  TRY(dev_upload(&d_v, v8.data(), n, w->stream));
  NewInfArgs a = make_newinf_args(w);
  a.member = d_p; a.variant = d_v; a.count_dev = nullptr; a.count_host = (uint32_t)n; a.cap = (uint32_t)n;
  StepDev sd;
  memset(&sd, 0, sizeof(sd));
  sd.now = time; sd.seed = seed; sd.step = step; sd.tpar = w->trans_par;
  CK(cudaMemcpyAsync(w->d_step_aux, &sd, sizeof(sd), cudaMemcpyHostToDevice, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  a.sp = reinterpret_cast<const StepDev*>(w->d_step_aux);
  unsigned blocks = (unsigned)std::min<size_t>((n + 31) / 32, JB_NEWINF_GRID);
  k_newinf<<<blocks, JB_NEWINF_THREADS, 0, w->stream>>>(a, w->dis_host);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(w->stream));
  cudaFree(d_p);
  cudaFree(d_v);
  uint32_t hw = 0;
  CK(cudaMemcpy(&hw, w->counters + CNT_SLOTS_HW, 4, cudaMemcpyDeviceToHost));
  if (hw > w->C) {  // the infections beyond the table were dropped by the kernel: report it, keep the table consistent
    CK(cudaMemcpy(w->counters + CNT_SLOTS_HW, &w->C, 4, cudaMemcpyHostToDevice));
    jb_set_error("jb_infect: infection slot capacity exceeded (JbWorldDesc.slot_capacity)");
    return JB_ECAP;
  }
  return JB_OK;
}

extern "C" int jb_mutate(JbWorld* w, uint32_t variant, const double* regional_probability, uint64_t seed, uint32_t step) {
  if (!w || !regional_probability) { jb_set_error("null argument"); return JB_EINVAL; }
  REQUIRE(variant < w->V, "jb_mutate: variant out of range");
  REQUIRE(!w->in_step, "jb_mutate inside a step");
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  uint32_t hw = 0;
  CK(cudaMemcpy(&hw, w->counters + CNT_SLOTS_HW, 4, cudaMemcpyDeviceToHost));
  if (!hw) return JB_OK;
  double* d_prob = nullptr;
  TRY(dev_upload(&d_prob, regional_probability, w->R, w->stream));
  CK(cudaMemsetAsync(w->counters + CNT_NEWINF, 0, 4, w->stream));
  k_mutate_select<<<(hw + 255) / 256, 256, 0, w->stream>>>(w->inf_person, w->inf_region, hw, d_prob, w->int2ext, (uint32_t)w->id_base,
                                                           seed, step, w->newinf_member, w->newinf_cap, w->counters + CNT_NEWINF);
  CK(cudaGetLastError());
  NewInfArgs a = make_newinf_args(w);
  a.member = w->newinf_member; a.variant = nullptr; a.count_dev = w->counters + CNT_NEWINF; a.cap = w->newinf_cap;
  a.mutate = 1; a.mutate_variant = variant;
  StepDev sd;
  memset(&sd, 0, sizeof(sd));
  sd.seed = seed; sd.step = step; sd.tpar = w->trans_par;
  CK(cudaMemcpyAsync(w->d_step_aux, &sd, sizeof(sd), cudaMemcpyHostToDevice, w->stream));
  a.sp = reinterpret_cast<const StepDev*>(w->d_step_aux);
  k_newinf<<<JB_NEWINF_GRID, JB_NEWINF_THREADS, 0, w->stream>>>(a, w->dis_host);
  CK(cudaGetLastError());
  CK(cudaMemsetAsync(w->counters + CNT_NEWINF, 0, 4, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  cudaFree(d_prob);
  return JB_OK;
}

extern "C" int jb_upload_infections(JbWorld* w, const JbInfectionRow* rows, size_t n) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  if (n == 0) return JB_OK;
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  uint32_t hw = 0;
  CK(cudaMemcpy(&hw, w->counters + CNT_SLOTS_HW, 4, cudaMemcpyDeviceToHost));
  REQUIRE((size_t)hw + n <= w->C, "jb_upload_infections: slot capacity exceeded");
  const size_t C = w->C;
  // Host-side scatter through small staging vectors (cold path: seeding / checkpoint restore / tests).
  std::vector<int32_t> person(n);
  std::vector<double> start(n), next(n), par(n * JB_N_TPAR), tt(n * JB_MAX_STAGES);
  std::vector<uint8_t> stage(n), nst(n), var(n);
  std::vector<int8_t> tag(n), maxtag(n), ttag(n * JB_MAX_STAGES);
  for (size_t i = 0; i < n; ++i) {
    const JbInfectionRow& r = rows[i];
    REQUIRE(r.person < w->N_ext, "jb_upload_infections: person out of range");
    REQUIRE(r.n_stages >= 1 && r.n_stages <= JB_MAX_STAGES && r.stage >= 0 && r.stage < r.n_stages, "bad stages");
    REQUIRE(r.variant < w->V, "bad variant");
    person[i] = (int32_t)w->h_ext2int[r.person]; start[i] = r.start_time;
    stage[i] = (uint8_t)r.stage; nst[i] = (uint8_t)r.n_stages; var[i] = (uint8_t)r.variant;
    tag[i] = (int8_t)r.tag; maxtag[i] = (int8_t)r.max_tag;
    next[i] = (r.stage + 1 < r.n_stages) ? r.traj_time[r.stage + 1] : 1e300;
    for (int q = 0; q < JB_N_TPAR; ++q) par[(size_t)q * n + i] = r.tpar[q];
    for (int q = 0; q < JB_MAX_STAGES; ++q) {
      tt[(size_t)q * n + i] = q < r.n_stages ? r.traj_time[q] : 1e300;
      ttag[(size_t)q * n + i] = q < r.n_stages ? (int8_t)r.traj_tag[q] : 0;
    }
  }
  CK(cudaMemcpy(w->inf_person + hw, person.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(w->inf_start + hw, start.data(), n * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(w->inf_next + hw, next.data(), n * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(w->inf_stage + hw, stage.data(), n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(w->inf_nst + hw, nst.data(), n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(w->inf_var + hw, var.data(), n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(w->inf_tag + hw, tag.data(), n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(w->inf_maxtag + hw, maxtag.data(), n, cudaMemcpyHostToDevice));
  {
    std::vector<uint8_t> reg(n), sev(n);
    std::vector<int32_t> hosp(n), med(n);
    for (size_t i = 0; i < n; ++i) {
      const uint32_t pi = (uint32_t)person[i];
      reg[i] = w->h_pregion[pi]; hosp[i] = w->h_phosp[pi];
      sev[i] = (JB_TAG_BIT(rows[i].tag) & w->dis_host.severe_mask) ? 1 : 0;
    }
    CK(cudaMemcpy(w->inf_region + hw, reg.data(), n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w->inf_sev + hw, sev.data(), n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w->inf_hosp + hw, hosp.data(), n * 4, cudaMemcpyHostToDevice));
    (void)med;  // (the slot's ward mirror is gathered from the person's ward by k_upload_person_side)
  }
  for (int q = 0; q < JB_N_TPAR; ++q)
    CK(cudaMemcpy(w->inf_par + (size_t)q * C + hw, par.data() + (size_t)q * n, n * 8, cudaMemcpyHostToDevice));
  if (w->dis_host.trans_type == JB_TRANS_GAMMA) {  // the derived row of the gamma profile (jb_gamma_front)
    k_gamma_front<<<(unsigned)((n + 255) / 256), 256, 0, w->stream>>>(w->inf_par, C, hw, (uint32_t)n);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(w->stream));
  }
  for (int q = 0; q < JB_MAX_STAGES; ++q) {
    CK(cudaMemcpy(w->traj_time + (size_t)q * C + hw, tt.data() + (size_t)q * n, n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w->traj_tag + (size_t)q * C + hw, ttag.data() + (size_t)q * n, n, cudaMemcpyHostToDevice));
  }
  // per-person side (state byte, slot, tag, variant, transmission probability, cross immunity, the slot's ward mirror):
  // one scatter kernel over the staged rows instead of a dozen small copies per row (a checkpoint of a large world
  // restores millions of rows)
  {
    std::vector<double> prob(n);
    for (size_t i = 0; i < n; ++i) prob[i] = rows[i].probability;
    int32_t* d_person = nullptr; int8_t* d_tag = nullptr; uint8_t* d_var = nullptr; double* d_prob = nullptr;
    TRY(dev_upload(&d_person, person.data(), n, w->stream));
    TRY(dev_upload(&d_tag, tag.data(), n, w->stream));
    TRY(dev_upload(&d_var, var.data(), n, w->stream));
    TRY(dev_upload(&d_prob, prob.data(), n, w->stream));
    JbCrossMasks cm;
    for (int v = 0; v < JB_MAX_VARIANTS; ++v) cm.c[v] = w->dis_host.cross[v];
    k_upload_person_side<<<(unsigned)((n + 255) / 256), 256, 0, w->stream>>>(
        d_person, d_tag, d_var, d_prob, (uint32_t)n, hw, w->NP, (int)w->V, cm, w->dis_host.severe_mask, w->pstate, w->pslot, w->ptag,
        w->pvar, w->trans[0], w->trans[1], w->susc, w->pmed, w->inf_med);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(w->stream));
    cudaFree(d_person); cudaFree(d_tag); cudaFree(d_var); cudaFree(d_prob);
  }
  hw += (uint32_t)n;
  CK(cudaMemcpy(w->counters + CNT_SLOTS_HW, &hw, 4, cudaMemcpyHostToDevice));
  return JB_OK;
}

extern "C" int jb_set_halo(JbWorld* w, const uint32_t* out_person, size_t n_out) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  std::vector<uint32_t> op(n_out);
  for (size_t i = 0; i < n_out; ++i) {
    REQUIRE(out_person[i] < w->N_ext, "jb_set_halo: person out of range");
    op[i] = w->h_ext2int[out_person[i]];
  }
  if (w->out_person) cudaFree(w->out_person);
  TRY(dev_upload(&w->out_person, op.data(), n_out, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  w->n_out = (uint32_t)n_out;
  return JB_OK;
}

extern "C" int jb_set_halo_gids(JbWorld* w, const uint32_t* gids, size_t n) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  REQUIRE(n == w->H, "jb_set_halo_gids: need n_halo entries");
  CK(cudaSetDevice(w->device));
  if (n) CK(cudaMemcpy(w->halo_gid, gids, n * 4, cudaMemcpyHostToDevice));
  if (n && w->mgid) {  // visitors registered in a mirrored supergroup: their id goes to their slot as well
    k_mirror_halo_gids<<<(unsigned)((n + 255) / 256), 256, 0, w->stream>>>(w->pmslot + w->N, w->halo_gid, (uint32_t)n, w->mgid);
    KCHECK("k_mirror_halo_gids", w->stream);
    CK(cudaStreamSynchronize(w->stream));
  }
  return JB_OK;
}

// ---- the step ----------------------------------------------------------------------------
static GroupArgs make_group_args(JbWorld* w, const JbSupergroup& sg, int L) {
  GroupArgs a;
  memset(&a, 0, sizeof(a));
  const JbStepParams& p = w->cur;
  a.grp_off = w->grp_off; a.grp_info = w->grp_info; a.grp_aux = w->grp_aux;
  a.reg_member = w->reg_member; a.reg_info = w->reg_info;
  a.phot = w->phot; a.susc = w->susc; a.trans = w->trans[0]; a.halo_gid = w->halo_gid; a.mgid = w->mgid;
  a.school_years = w->school_years; a.cm = w->cm + sg.cm_offset; a.beta = w->beta; a.beta_scale = w->beta_scale;
  a.dyn_cnt = sg.has_dynamic ? w->dyn_cnt : nullptr;
  const size_t ksg = &sg - w->sgs.data();
  a.dyn_bin = sg.has_dynamic ? w->dyn_bin + w->sg_bin_slot[ksg] : nullptr;  // the supergroup's first bin
  a.bin_cap = w->sg_bin_cap[ksg];
  a.dyn_bin_base = w->sg_bin_base[ksg];
  a.g0 = sg.first_group; a.g1 = sg.first_group + sg.n_groups;
  a.group_list = nullptr; a.n_list = sg.n_groups; a.int2ext = w->int2ext; a.ext2int = w->ext2int;
  a.dim = sg.dim; a.kind = sg.kind; a.L = L;
  a.static_acts = w->sg_static_acts[&sg - w->sgs.data()];
  a.N = w->N; a.NP = w->NP; a.V = (int)w->V; a.R = (int)w->R;
  a.sp = step_dev(w); a.gid_base = (uint32_t)w->id_base;
  a.counters = w->counters; a.newinf_member = w->newinf_member; a.newinf_var = w->newinf_var;
  a.newinf_cap = w->newinf_cap;
  if (p.flags & JB_STEP_RECORD_EVENTS) {
    a.ev_infected = w->ev_infected; a.ev_infector = w->ev_infector; a.ev_group = w->ev_group; a.ev_var = w->ev_var;
  }
  a.lambda = (p.flags & JB_STEP_RECORD_LAMBDA) ? w->lambda : nullptr;
  if (p.flags & JB_STEP_INJECT_UNIFORMS) {
    a.inject = w->inject; a.draw_base = w->draw_base;
  }
  a.draw_cnt = w->draw_cnt;
  a.mode = 0;
  return a;
}

template <int VM, int WPG>
static int launch_list(JbWorld* w, const GroupArgs& a, bool dynamic, cudaStream_t st) {
  if (a.n_list == 0) return JB_OK;
  constexpr int gpc = JB_CTA_WARPS / WPG;
  const size_t smem = jb_grp_smem_bytes(a.L, (int)w->V, dynamic ? a.bin_cap : 0u, WPG) * gpc;
  if (a.ev_group) k_groups<VM, WPG, true><<<(a.n_list + gpc - 1) / gpc, JB_CTA, smem, st>>>(a);
  else k_groups<VM, WPG, false><<<(a.n_list + gpc - 1) / gpc, JB_CTA, smem, st>>>(a);
  KCHECK(WPG == 1 ? "k_groups<warp>" : "k_groups<cta>", st);
  return JB_OK;
}

template <int VM>
static int launch_span(JbWorld* w, const GroupArgs& a, cudaStream_t st) {
  constexpr unsigned warps = JB_SPAN_WARPS_OF(VM);
  const unsigned blocks = (a.n_span + warps - 1) / warps;
  if (a.ev_group) k_span<VM, true><<<blocks, 32 * warps, 0, st>>>(a);
  else k_span<VM, false><<<blocks, 32 * warps, 0, st>>>(a);
  KCHECK("k_span", st);
  return JB_OK;
}

template <int TM, int VM>
static int launch_tiles(JbWorld* w, const GroupArgs& a, cudaStream_t st) {
  const unsigned blocks = (a.n_tiles * 2 + JB_TILE_THREADS - 1) / JB_TILE_THREADS;
  if (a.ev_group) k_tiles<TM, VM, true><<<blocks, JB_TILE_THREADS, 0, st>>>(a);
  else k_tiles<TM, VM, false><<<blocks, JB_TILE_THREADS, 0, st>>>(a);
  KCHECK(TM == JB_TM_STREAM ? "k_tiles<stream>" : TM == JB_TM_MIRROR ? "k_tiles<mirror>" : "k_tiles<gather>", st);
  return JB_OK;
}

// Individual policies only ever REMOVE activities other than medical_facility and residence (stay-home policies keep
// exactly those two, skip-activity policies drop primary_activity / commute: individual_policies.py:87-119,441-824).
// In a step that offers nothing else they cannot change anybody's placement, so the table-driven placement is exact.
static inline bool policies_decide_per_person(const JbWorld* w, uint32_t activity_mask) {
  const JbStepParams& p = w->cur;
  uint32_t other = activity_mask & ~((1u << JB_ACT_MEDICAL) | (1u << JB_ACT_RESIDENCE)) & ((1u << JB_N_ACT) - 1u);
  // (an activity nobody can take in this world / step is not an offer: commute without transports, leisure without tables)
  if (!(w->cmt.person != nullptr || ((p.flags & JB_STEP_INJECT_COMMUTE) && w->cm_inject != nullptr))) other &= ~(1u << JB_ACT_COMMUTE);
  if (!(w->lz.K != 0 && (p.leisure_table != 0 || (p.flags & JB_STEP_INJECT_LEISURE)))) other &= ~(1u << JB_ACT_LEISURE);
  return w->n_pol != 0 && !w->pol_subject_mode && other != 0u;
}

// Does this step draw leisure activities with residence visits among them?  Then the leisure-goers are placed in three
// moves (k_place records the destinations, k_visit_* resolve who is kept home, the rest are scattered).
static inline bool visits_step(const JbWorld* w) {
  const JbStepParams& p = w->cur;
  return w->vz.activity >= 0 && w->lz.K != 0 && (p.leisure_table != 0 || (p.flags & JB_STEP_INJECT_LEISURE)) &&
         (p.activity_mask & (1u << JB_ACT_LEISURE)) != 0u;
}

// Does this step place people one by one (individual policies, leisure or commute draws)?  Then the mirrored
// supergroups read the per-step copy the placement writes instead of the persistent mirror.
static inline bool step_places_per_person(const JbWorld* w) {
  const JbStepParams& p = w->cur;
  const bool commute = (p.activity_mask & (1u << JB_ACT_COMMUTE)) &&
                       (w->cmt.person != nullptr || ((p.flags & JB_STEP_INJECT_COMMUTE) && w->cm_inject != nullptr));
  return policies_decide_per_person(w, p.activity_mask) || (w->lz.K != 0 && (p.leisure_table != 0 || (p.flags & JB_STEP_INJECT_LEISURE))) || commute;
}

static int launch_groups(JbWorld* w, int mode) {
  // fork: every active supergroup gets its own streams (its groups are disjoint from all others,
  // person state is read-only here, outputs are appended with atomics); inside a supergroup the
  // tiles, the warp list and the CTA list are independent launches as well
  CK(cudaEventRecord(w->ev_fork, w->stream));
  // Kernels that become ready together are started in the order their nodes were created, and a kernel whose blocks fill
  // the SMs keeps the ones behind it waiting until it drains.  On a primary-activity step (companies, schools and
  // households: three long kernels) the supergroups are launched smallest first, by tile / list sizes, so that the short
  // kernels do not queue behind the long ones: 906 -> 867 us at 56 M people, 186 -> 168 us at 7 M.  On the other steps
  // the supergroup order (hospitals, households, care homes) measured 1-4 % faster than smallest-first, and is kept.
  // JB_LAUNCH_ORDER=0: supergroup order always; 2: smallest first always.
  static const int order_mode = jb_env_flag("JB_LAUNCH_ORDER", 1);
  const bool by_size = order_mode == 2 || (order_mode == 1 && (w->cur.activity_mask & (1u << JB_ACT_PRIMARY)) != 0u);
  std::vector<size_t> order(w->sgs.size());
  for (size_t k = 0; k < order.size(); ++k) order[k] = k;
  if (by_size && !w->serial) {
    auto weight = [&](size_t k) -> uint64_t {
      const SgTiles& T = w->tiles[k];
      return (uint64_t)T.n_tiles * 32u + (uint64_t)T.n_span * 512u + (uint64_t)T.n_w * 64u + (uint64_t)T.n_c * 256u;
    };
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return weight(x) < weight(y); });
  }
  for (size_t oi = 0; oi < order.size(); ++oi) {
    const size_t k = order[oi];
    const JbSupergroup& sg = w->sgs[k];
    if (sg.n_groups == 0) continue;
    if (!(w->cur.activity_mask & (1u << sg.activity))) continue;
    GroupArgs a = make_group_args(w, sg, w->sg_L[k]);
    a.mode = mode;
    const bool tme = w->timing >= 2 && mode == 0;
    if (tme) CK(cudaEventRecord(w->ev_sg0[k], w->stream));
    const SgTiles& T = w->tiles[k];
    const bool multi = w->V > 1;
    if (T.mode == 3) { a.marr = step_places_per_person(w) ? w->mstep : w->mstate; a.stream_base = T.stream_base; }
    const bool visited_here = (int)k == w->visits_sg && visits_step(w);
    if (visited_here) a.vslot = w->vz.vslot;  // households that are being visited are left to the visited-list launch
    for (int u = 0; u < 3; ++u) {
      if ((u == 0 && !T.n_span && (T.mode == 0 || T.n_tiles == 0 || !T.has_small)) || (u == 1 && T.n_w == 0) || (u == 2 && T.n_c == 0)) continue;
      static const int unit_streams = jb_env_flag("JB_UNIT_STREAMS", 1);
      cudaStream_t st = w->serial ? w->stream : w->side[k][unit_streams ? u : 0];
      if (!w->serial) CK(cudaStreamWaitEvent(st, w->ev_fork, 0));
      if (!w->serial && w->halo_async && w->sg_has_halo[k]) CK(cudaStreamWaitEvent(st, w->ev_halo_ready, 0));  // its visitors have arrived
      if (u == 0) {
        static const int skip_b = jb_env_flag("JB_TILES_SKIP_B", 0);
        if (skip_b && mode == 0) a.mode = 2;
        a.tile_group0 = T.tile_group0; a.tile_ginfo = T.tile_ginfo; a.tp_member = T.tp_member; a.tp_info = T.tp_info;
        static const int lean_on = jb_env_flag("JB_TILES_LEAN", 1);
        a.tile_static = lean_on ? T.tile_static : nullptr; a.uniform_act = T.uniform_act;
        a.n_tiles = T.n_tiles; a.stream_base = T.stream_base;
        a.span_items = T.span_items; a.n_span = T.n_span; a.span_group = T.span_group;
        if (T.n_span) { if (!multi) TRY((launch_span<1>(w, a, st))); else TRY((launch_span<JB_MAX_VARIANTS>(w, a, st))); }
        else if (T.mode == 2) { if (!multi) TRY((launch_tiles<JB_TM_STREAM, 1>(w, a, st))); else TRY((launch_tiles<JB_TM_STREAM, JB_MAX_VARIANTS>(w, a, st))); }
        else if (T.mode == 3) { if (!multi) TRY((launch_tiles<JB_TM_MIRROR, 1>(w, a, st))); else TRY((launch_tiles<JB_TM_MIRROR, JB_MAX_VARIANTS>(w, a, st))); }
        else { if (!multi) TRY((launch_tiles<JB_TM_GATHER, 1>(w, a, st))); else TRY((launch_tiles<JB_TM_GATHER, JB_MAX_VARIANTS>(w, a, st))); }
      } else if (u == 1) {
        a.group_list = T.list_w; a.n_list = T.n_w; a.list_slot = T.list_w_slot;
        if (!multi) TRY((launch_list<1, 1>(w, a, false, st))); else TRY((launch_list<JB_MAX_VARIANTS, 1>(w, a, false, st)));
      } else {
        a.group_list = T.list_c; a.n_list = T.n_c; a.list_slot = T.list_c_slot;
        if (!multi) TRY((launch_list<1, JB_CTA_WARPS>(w, a, sg.has_dynamic != 0, st)));
        else TRY((launch_list<JB_MAX_VARIANTS, JB_CTA_WARPS>(w, a, sg.has_dynamic != 0, st)));
      }
      w->launches++;
      if (!w->serial) {
        CK(cudaEventRecord(w->ev_join[k][u], st));
        CK(cudaStreamWaitEvent(w->stream, w->ev_join[k][u], 0));
      }
    }
    if (visited_here) {
      // the households being visited in this step, their visitors as dynamic members, beta rule "household_visits":
      // one warp per household, the list's length is on the device
      GroupArgs v = make_group_args(w, sg, w->sg_L[k]);
      v.mode = mode;
      v.group_list = w->vz.vis_group; v.n_list = w->vz.vis_cap; v.n_list_dev = w->counters + CNT_VIS_N;
      v.dyn_cnt = w->vz.vis_cnt; v.dyn_bin = w->vz.vis_bin; v.bin_cap = JB_VIS_BIN; v.dyn_bin_base = 0; v.dyn_by_list = 1;
      v.beta_rule = w->visits_beta + 1;
      cudaStream_t st = w->serial ? w->stream : w->side[k][1];
      if (!w->serial) CK(cudaStreamWaitEvent(st, w->ev_fork, 0));
      {
        constexpr int gpc = JB_CTA_WARPS;
        const size_t smem = jb_grp_smem_bytes(v.L, (int)w->V, v.bin_cap, 1) * gpc;
        const unsigned blocks = std::min<unsigned>((w->vz.vis_cap + gpc - 1) / gpc, 148u * 16u);
        if (!multi) { if (v.ev_group) k_groups<1, 1, true><<<blocks, JB_CTA, smem, st>>>(v); else k_groups<1, 1, false><<<blocks, JB_CTA, smem, st>>>(v); }
        else { if (v.ev_group) k_groups<JB_MAX_VARIANTS, 1, true><<<blocks, JB_CTA, smem, st>>>(v); else k_groups<JB_MAX_VARIANTS, 1, false><<<blocks, JB_CTA, smem, st>>>(v); }
        KCHECK("k_groups<visited households>", st);
      }
      w->launches++;
      if (!w->serial) {
        CK(cudaEventRecord(w->ev_join[k][1], st));
        CK(cudaStreamWaitEvent(w->stream, w->ev_join[k][1], 0));
      }
    }
    CK(cudaGetLastError());
    if (tme) { CK(cudaEventRecord(w->ev_sg1[k], w->stream)); w->sg_pending[k] = true; }
  }
  return JB_OK;
}

#define JB_EVENT_RING 512
static int ring_fold_oldest(JbWorld* w) {
  const size_t i = (size_t)(w->ring_tail % JB_EVENT_RING);
  CK(cudaEventSynchronize(w->ring1[i]));
  float b = 0;
  CK(cudaEventElapsedTime(&b, w->ring0[i], w->ring1[i]));
  w->acc_step_ms += b; w->acc_steps++;
  w->ring_tail++;
  return JB_OK;
}
// Start / stop of a step's event bracket.  Level 1: ring (no host wait unless it wraps); level 2: the
// single detailed pair, folded before the next step.
static int timing_begin(JbWorld* w, cudaStream_t s);
static int timing_end(JbWorld* w, cudaStream_t s);

// Fold finished event pairs into the accumulators (called with the stream idle or synchronised).
static int harvest_timing(JbWorld* w) {
  while (w->ring_tail < w->ring_head) TRY(ring_fold_oldest(w));
  if (!w->ev_pending) return JB_OK;
  CK(cudaEventSynchronize(w->ev_step1));
  float a = 0, b = 0;
  if (w->ev_detail) CK(cudaEventElapsedTime(&a, w->ev_int0, w->ev_int1));
  CK(cudaEventElapsedTime(&b, w->ev_step0, w->ev_step1));
  w->acc_int_ms += a; w->acc_step_ms += b; w->acc_steps++;
  for (size_t k = 0; k < w->sgs.size(); ++k)
    if (w->sg_pending[k]) {
      float c = 0;
      CK(cudaEventElapsedTime(&c, w->ev_sg0[k], w->ev_sg1[k]));
      w->acc_sg_ms[k] += c; w->acc_sg_launches[k]++;
      w->sg_pending[k] = false;
    }
  w->ev_pending = false;
  return JB_OK;
}

static int timing_begin(JbWorld* w, cudaStream_t s) {
  if (w->timing == 1) {
    if (w->ring0.empty()) {
      w->ring0.resize(JB_EVENT_RING); w->ring1.resize(JB_EVENT_RING);
      for (int i = 0; i < JB_EVENT_RING; ++i) { CK(cudaEventCreate(&w->ring0[i])); CK(cudaEventCreate(&w->ring1[i])); }
    }
    if (w->ring_head - w->ring_tail == JB_EVENT_RING) TRY(ring_fold_oldest(w));
    CK(cudaEventRecord(w->ring0[w->ring_head % JB_EVENT_RING], s));
  } else if (w->timing) {
    TRY(harvest_timing(w));
    CK(cudaEventRecord(w->ev_step0, s));
  }
  return JB_OK;
}
static int timing_end(JbWorld* w, cudaStream_t s) {
  if (w->timing == 1) {
    CK(cudaEventRecord(w->ring1[w->ring_head % JB_EVENT_RING], s));
    w->ring_head++;
  } else if (w->timing) {
    CK(cudaEventRecord(w->ev_step1, s));
    w->ev_pending = true; w->ev_detail = w->timing >= 2;
  }
  return JB_OK;
}

// ---- device work of the three step phases (no host synchronisation: capturable) ----------

static inline bool halo_step(const JbWorld* w);
static int enqueue_begin(JbWorld* w, void* d_send) {
  cudaStream_t s = w->stream;
  const JbStepParams& p = w->cur;
  {
    const uint32_t n_zero = (uint32_t)(w->counter_words - CNT_STEP_BEGIN);
    k_step_begin<<<std::max(1u, std::min(148u, (n_zero + 1023) / 1024)), 256, 0, s>>>(
        w->d_step_ring, (uint32_t)(w->step_stride / 4), reinterpret_cast<uint32_t*>(w->d_step), (uint32_t)((w->step_bytes + 3) / 4),
        w->d_step_seq, w->counters + CNT_STEP_BEGIN, n_zero);
    KCHECK("k_step_begin", s);
    w->launches++;
  }
  if (p.flags & JB_STEP_RECORD_LAMBDA) {
    k_fill_f64<<<(w->NP + 255) / 256, 256, 0, s>>>(w->lambda, w->NP, std::numeric_limits<double>::quiet_NaN());
    w->launches++;
  }
  if (w->N) {
    PlaceArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.pstate = w->pstate; pa.phot = w->phot; pa.pvar = w->V > 1 ? w->pvar : nullptr; pa.phas = w->phas; pa.pmed = w->pmed; pa.dyn_cnt = w->dyn_cnt; pa.dyn_bin = w->dyn_bin;
    pa.counters = w->counters; pa.N = w->N; pa.sp = step_dev(w);
    pa.dr = w->dyn_ranges;
    pa.lz = w->lz;
    pa.lz.inject = (p.flags & JB_STEP_INJECT_LEISURE) ? w->lz_inject : nullptr;
    if ((p.flags & JB_STEP_INJECT_LEISURE) && !w->lz_inject) { jb_set_error("JB_STEP_INJECT_LEISURE without jb_inject_leisure"); return JB_EINVAL; }
    pa.age = w->age; pa.sex = w->sex; pa.ch = w->ch; pa.pregion = w->pregion; pa.int2ext = w->int2ext;
    pa.gid_base = (uint32_t)w->id_base; pa.R = (int)w->R;
    pa.cmt = w->cmt;
    pa.cmt.inject = (p.flags & JB_STEP_INJECT_COMMUTE) ? w->cm_inject : nullptr;
    if (!(p.activity_mask & (1u << JB_ACT_COMMUTE))) { pa.cmt.person = nullptr; pa.cmt.inject = nullptr; }
    pa.pol = w->d_pol; pa.n_pol = w->n_pol; pa.ptag = w->ptag; pa.pattr = w->pattr; pa.pwork_region = w->pwork_region;
    pa.pol_inject = (p.flags & JB_STEP_INJECT_POLICY) ? w->pol_inject : nullptr;
    pa.severe_mask = w->dis_host.severe_mask;
    pa.grp_info = w->grp_info; pa.summary = w->summary;
    pa.vz = w->vz; pa.vis_defer = visits_step(w) ? 1 : 0;
    const uint32_t chunks = (w->N + 15) / 16;
    const bool pol_now = policies_decide_per_person(w, p.activity_mask);
    if (!pol_now && !w->pol_subject_mode) { pa.pol = nullptr; pa.n_pol = 0; }  // (the subject list is re-evaluated by k_place_subjects)
    const bool full = pol_now || (w->lz.K != 0 && (p.leisure_table != 0 || (p.flags & JB_STEP_INJECT_LEISURE))) ||
                      pa.cmt.person != nullptr || pa.cmt.inject != nullptr;
    if (full != step_places_per_person(w)) { jb_set_error("internal: placement flavour mismatch"); return JB_EINVAL; }
    if (w->mir_slots && (p.activity_mask & (1u << JB_ACT_PRIMARY))) {  // the mirrored supergroups run in this step
      pa.pmslot = w->pmslot;
      if (full) pa.mstep = w->mstep;
      else { pa.mstate = w->mstate; pa.pmir = w->pmir; }
    }
    if (full) {
      // a grid of a few waves, not one block per 256 people: every block ends with five atomics on the same five
      // per-activity counters (a million same-address atomics at 56 M people: 12 % of the kernel's stall samples)
      // and starts by building the 512-entry activity table
      static const uint32_t full_ctas = (uint32_t)std::max(1, jb_env_flag("JB_PLACE_FULL_CTAS", 12));
      k_place<true><<<std::min<uint32_t>((w->N + 255) / 256, 148 * full_ctas), 256, 0, s>>>(pa);
    }
    else {
      // (six CTAs of <= 42 registers are resident per SM; a finer grid than one wave measured faster: blocks that finish
      // early are replaced while the others are still streaming)
      static const uint32_t ctas = (uint32_t)std::max(1, jb_env_flag("JB_PLACE_CTAS", 16));
      k_place<false><<<std::min<uint32_t>((chunks + 255) / 256, 148 * ctas), 256, 0, s>>>(pa);
    }
    KCHECK("k_place", s);
    w->launches++;
    if (!full && w->pol_subject_mode && w->n_pol_subjects) {
      k_place_subjects<<<(w->n_pol_subjects + 255) / 256, 256, 0, s>>>(pa, w->pol_subjects, w->n_pol_subjects);
      KCHECK("k_place_subjects", s);
      w->launches++;
    }
    if (pa.vis_defer) {
      if (!full) { jb_set_error("internal: residence visits need the per-person placement"); return JB_EINVAL; }
      // who is kept home by a visitor: JB_VIS_ITERATIONS passes over the people who drew a destination (buffers rotate:
      // pass k reads k - 1 mod 3, accumulates into k mod 3, clears k + 1 mod 3), then everybody is placed
      const unsigned vb = 148 * 4;
      for (int k = 1; k <= JB_VIS_ITERATIONS; ++k) {
        k_visit_iter<<<vb, 256, 0, s>>>(w->vz, w->counters, w->int2ext, (k - 1) % 3, k % 3, (k + 1) % 3);
        KCHECK("k_visit_iter", s);
      }
      k_visit_final<<<vb, 256, 0, s>>>(pa, JB_VIS_ITERATIONS % 3);
      KCHECK("k_visit_final", s);
      k_visit_scatter<<<vb, 256, 0, s>>>(w->vz, w->counters, w->int2ext);
      KCHECK("k_visit_scatter", s);
      w->launches += JB_VIS_ITERATIONS + 2;
    }
  }
  if (w->p2p_on && halo_step(w)) {
    k_halo_push<<<std::max(1u, (w->n_out + 255) / 256), 256, 0, s>>>(w->p2p, w->out_person, w->n_out, w->phot, w->susc, w->trans[0],
                                                                    w->NP, (int)w->V, step_dev(w), w->p2p_done);
    KCHECK("k_halo_push", s);
    w->launches++;
  } else if (w->n_out && d_send && halo_step(w)) {
    k_halo_pack<<<(w->n_out + 255) / 256, 256, 0, s>>>(w->out_person, w->n_out, w->phot, w->susc, w->trans[0],
                                                       w->NP, (int)w->V, (unsigned char*)d_send, step_dev(w));
    KCHECK("k_halo_pack", s);
    w->launches++;
  }
  return JB_OK;
}

// The infections created in a step get their first health update from k_newinf itself, so
// k_health only walks the slots that existed at step start.  Single-domain and peer-memory steps
// launch it beside the interaction kernels (enqueue_interact); residents infected abroad arrive
// later in the step and get their first update from k_newinf as well.  (With host-issued
// collectives the step is three separately captured phases and k_health runs in the last one; when
// the host creates the infections, JB_STEP_NO_NEW_INFECTIONS, k_health walks every slot there.)
// Does this step exchange halo persons?  Only when one of the activities under which halo persons are
// registered (in any domain) takes place: a commuter at home is present in no group of the other domain,
// and the supergroups that hold halo registrations are not even processed.
static inline bool halo_step_for(const JbWorld* w, uint32_t activity_mask) {
  return (w->H || w->n_out) && (activity_mask & w->halo_act_mask) != 0u;
}
static inline bool halo_step(const JbWorld* w) { return halo_step_for(w, w->cur.activity_mask); }

static inline bool fused_finish(const JbWorld* w) {
  return (!halo_step(w) || w->p2p_on) && !(w->cur.flags & JB_STEP_NO_NEW_INFECTIONS);
}

static HealthArgs make_health_args(JbWorld* w, int old_only) {
  HealthArgs a;
  memset(&a, 0, sizeof(a));
  a.C = w->C; a.sp = step_dev(w); a.dis = w->dis; a.old_only = old_only;
  a.masks.trans_type = w->dis_host.trans_type; a.masks.rec_mask = w->dis_host.rec_mask;
  a.masks.dead_mask = w->dis_host.dead_mask; a.masks.hosp_mask = w->dis_host.hosp_mask;
  a.masks.icu_mask = w->dis_host.icu_mask; a.masks.severe_mask = w->dis_host.severe_mask;
  a.inf_person = w->inf_person; a.inf_start = w->inf_start; a.inf_par = w->inf_par; a.inf_next = w->inf_next;
  a.traj_time = w->traj_time; a.inf_stage = w->inf_stage; a.inf_nst = w->inf_nst; a.inf_tag = w->inf_tag;
  a.traj_tag = w->traj_tag; a.pstate = w->pstate; a.ptag = w->ptag; a.trans = w->trans[0]; a.NP = w->NP; a.pmed = w->pmed;
  a.pslot = w->pslot; a.phas = w->phas; a.counters = w->counters;
  a.inf_region = w->inf_region; a.inf_hosp = w->inf_hosp; a.inf_med = w->inf_med; a.inf_sev = w->inf_sev;
  a.summary = w->summary; a.R = (int)w->R; a.grp_info = w->grp_info;
  if ((w->cur.flags & (JB_STEP_RECORD_EVENTS | JB_STEP_RECORD_HEALTH)) && w->hev_person) {
    a.hev_person = w->hev_person; a.hev_kind = w->hev_kind; a.hev_group = w->hev_group; a.hev_cap = w->hev_cap;
    a.pres_group = w->pres_group;
  }
  return a;
}

static int launch_health(JbWorld* w, cudaStream_t s) {
  const HealthArgs a = make_health_args(w, (w->cur.flags & JB_STEP_NO_NEW_INFECTIONS) ? 0 : 1);
  k_health<<<148 * 4, 256, 0, s>>>(a);
  KCHECK("k_health", s);
  w->launches++;
  return JB_OK;
}

static int enqueue_interact(JbWorld* w, const void* d_recv, void* d_ret_send) {
  cudaStream_t s = w->stream;
  const JbStepParams& p = w->cur;
  const bool halo_now = halo_step(w);
  // The arrival of the visitors (spin on the peers' flags, unpack) runs on its own stream: only the supergroups that
  // hold visitor registrations wait for it (launch_groups), the others start right behind the placement.
  static const int halo_async_on = jb_env_flag("JB_HALO_ASYNC", 1);
  w->halo_async = halo_async_on && !w->serial && halo_now && w->H && (w->p2p_on || d_recv) &&
                  !(p.flags & JB_STEP_INJECT_UNIFORMS);
  cudaStream_t hs = w->halo_async ? w->side_halo : s;
  if (w->halo_async) { CK(cudaEventRecord(w->ev_fork_halo, s)); CK(cudaStreamWaitEvent(hs, w->ev_fork_halo, 0)); }
  if (w->p2p_on && halo_now) {
    d_recv = w->p2p.recv[w->p2p.rank];
    d_ret_send = w->p2p_ret_local;
    k_halo_wait<<<1, 32, 0, hs>>>(w->p2p.flag_people[w->p2p.rank], w->p2p, 0, step_dev(w), w->counters, w->p2p_done + 2);
    KCHECK("k_halo_wait", hs);
    w->launches++;
  }
  if (w->H && d_recv && halo_now) {
    k_halo_unpack<<<(w->H + 255) / 256, 256, 0, hs>>>((const unsigned char*)d_recv, w->H, w->N, w->NP, (int)w->V,
                                                     w->phot, w->susc, w->trans[0], (uint8_t*)d_ret_send, step_dev(w),
                                                     w->pmslot, w->mir_slots ? (step_places_per_person(w) ? w->mstep : w->mstate) : nullptr);
    KCHECK("k_halo_unpack", hs);
    w->launches++;
  }
  if (w->halo_async) CK(cudaEventRecord(w->ev_halo_ready, hs));
  if (w->timing >= 2) CK(cudaEventRecord(w->ev_int0, s));
  if (p.flags & JB_STEP_INJECT_UNIFORMS) {
    CK(cudaMemsetAsync(w->draw_cnt, 0, ((size_t)w->G + 1) * 4, s));
    TRY(launch_groups(w, 1));
    size_t tmp = w->cub_tmp_bytes;
    CK(cub::DeviceScan::ExclusiveSum(w->cub_tmp, tmp, w->draw_cnt, w->draw_base, (int)(w->G + 1), s));
    w->launches += 2;
  }
  // The health update of the infections that existed at step start reads nothing the interaction
  // writes and writes nothing it reads (transmission probabilities are double buffered), so it runs
  // beside the interaction kernels instead of behind them.
  const bool health_beside = fused_finish(w);
  static const int health_early = jb_env_flag("JB_HEALTH_EARLY", 1);
  static const int health_blocks = jb_env_flag("JB_HEALTH_BLOCKS", 4);
  auto fork_health = [&](int phase) -> int {  // phase 1: dependency point only; 2: launch only; 3: both
    cudaStream_t hs = w->serial ? s : w->side_health;
    if (!w->serial && (phase & 1)) CK(cudaEventRecord(w->ev_fork_health, s));
    if (!(phase & 2)) return JB_OK;
    if (!w->serial) CK(cudaStreamWaitEvent(hs, w->ev_fork_health, 0));
    const HealthArgs ha = make_health_args(w, 1);
    k_health<<<148 * health_blocks, 256, 0, hs>>>(ha);
    KCHECK("k_health", hs);
    w->launches++;
    if (!w->serial) CK(cudaEventRecord(w->ev_join_health, hs));
    return JB_OK;
  };
  // the kernels that become ready together are started in the order their nodes were created: the
  // interaction kernels (the critical path) first, the health update (which has slack) last
  if (health_beside && health_early == 1) TRY(fork_health(3));
  if (health_beside && health_early == 2) TRY(fork_health(1));
  TRY(launch_groups(w, 0));
  if (health_beside && health_early == 2) TRY(fork_health(2));
  if (w->halo_async) CK(cudaStreamWaitEvent(s, w->ev_halo_ready, 0));  // join (also when no active supergroup holds visitors)
  if (w->timing >= 2) CK(cudaEventRecord(w->ev_int1, s));
  // create infections for residents infected here; flag halo persons for their home domain
  if (!(p.flags & JB_STEP_NO_NEW_INFECTIONS) || (w->H && halo_now)) {
    NewInfArgs a = make_newinf_args(w);
    a.member = w->newinf_member; a.variant = w->newinf_var; a.count_dev = w->counters + CNT_NEWINF;
    a.cap = w->newinf_cap; a.sp = step_dev(w);
    a.halo_ret = (uint8_t*)d_ret_send;
    a.fuse_health = 1; a.h = make_health_args(w, 0);
    if (health_beside && !health_early) TRY(fork_health(3));
    k_newinf<<<JB_NEWINF_GRID, JB_NEWINF_THREADS, 0, s>>>(a, w->dis_host);
    KCHECK("k_newinf", s);
    w->launches++;
  }
  if (health_beside && !w->serial) CK(cudaStreamWaitEvent(s, w->ev_join_health, 0));
  if (w->p2p_on && halo_now) {
    k_halo_ret_push<<<std::max(1u, (w->H + 255) / 256), 256, 0, s>>>(w->p2p, w->p2p_ret_local, w->H, step_dev(w), w->p2p_done + 1);
    KCHECK("k_halo_ret_push", s);
    w->launches++;
  }
  return JB_OK;
}

static int enqueue_finish(JbWorld* w, const void* d_ret_recv) {
  cudaStream_t s = w->stream;
  const JbStepParams& p = w->cur;
  const bool halo_now = halo_step(w);
  if (w->p2p_on && halo_now) {
    d_ret_recv = w->p2p.ret[w->p2p.rank];
    k_halo_wait<<<1, 32, 0, s>>>(w->p2p.flag_ret[w->p2p.rank], w->p2p, 1, step_dev(w), w->counters, w->p2p_done + 2);
    KCHECK("k_halo_wait(ret)", s);
    w->launches++;
  }
  if (w->n_out && d_ret_recv && halo_now) {
    // residents infected abroad: reuse the new-infection list
    // (tell_domains_to_infect, epidemiology.py:359-401)
    CK(cudaMemsetAsync(w->counters + CNT_HALO_INF, 0, 4, s));
    k_halo_return<<<(w->n_out + 255) / 256, 256, 0, s>>>((const uint8_t*)d_ret_recv, w->out_person, w->n_out,
                                                         w->newinf_member, w->newinf_var, w->newinf_cap,
                                                         w->counters + CNT_NEWINF, w->counters + CNT_HALO_INF);
    KCHECK("k_halo_return", s);
    w->launches++;
    if (!(p.flags & JB_STEP_NO_NEW_INFECTIONS)) {
      NewInfArgs a = make_newinf_args(w);
      a.member = w->newinf_member; a.variant = w->newinf_var; a.count_dev = w->counters + CNT_HALO_INF;
      a.first_dev = w->counters + CNT_NEWINF;  // behind the step's own new infections
      a.cap = w->newinf_cap; a.sp = step_dev(w);
      a.fuse_health = 1; a.h = make_health_args(w, 0);
      k_newinf<<<148, JB_NEWINF_THREADS, 0, s>>>(a, w->dis_host);
      KCHECK("k_newinf(halo)", s);
      w->launches++;
    }
  }
  if (!fused_finish(w)) TRY(launch_health(w, s));
  if (visits_step(w)) {
    k_visit_cleanup<<<148 * 4, 256, 0, s>>>(w->vz, w->counters);
    KCHECK("k_visit_cleanup", s);
    w->launches++;
  }
  k_step_end<<<1, 256, 0, s>>>(w->counters, (uint32_t)(CNT_N + (size_t)w->R * JB_N_SUMMARY), w->d_h_counters);
  KCHECK("k_step_end", s);
  w->launches++;
  return JB_OK;
}

// Step graphs: every phase is captured once per key and replayed.  Graphs are keyed by everything that
// shapes the launch sequence (phase, activity mask, flags, exchange buffers); the per-step scalars --
// including which half of the transmission double buffer is read -- live in the device parameter block,
// so a replay needs no patching and one graph serves every step with the same activities.
static inline bool plain_launch(const JbWorld* w) {
  return !w->use_graphs || w->timing >= 2 || jb_debug_sync() ||
         (w->cur.flags & (JB_STEP_INJECT_UNIFORMS | JB_STEP_RECORD_LAMBDA));
}

// Index of the phase's graph in w->graphs, capturing and instantiating it first if it does not exist yet
// (nothing is executed here: the capture happens OUTSIDE the step's timing bracket).
template <typename F>
static int ensure_graph(JbWorld* w, int phase, const void* p0, const void* p1, F&& enqueue, size_t* index) {
  const JbStepParams& p = w->cur;
  const uint64_t k0 = ((uint64_t)phase << 56) | ((uint64_t)(w->serial ? 1 : 0) << 48) | ((uint64_t)(p.leisure_table ? 1 : 0) << 49) |
                      ((uint64_t)p.activity_mask << 32) | p.flags;
  const uint64_t k1 = (uint64_t)(uintptr_t)p0, k2 = (uint64_t)(uintptr_t)p1;
  for (size_t i = 0; i < w->graphs.size(); ++i)
    if (w->graphs[i].k0 == k0 && w->graphs[i].k1 == k1 && w->graphs[i].k2 == k2) { *index = i; return JB_OK; }
  const uint64_t before = w->launches;
  cudaGraph_t graph = nullptr;
  CK(cudaStreamBeginCapture(w->stream, cudaStreamCaptureModeThreadLocal));
  int rc = enqueue();
  cudaError_t ce = cudaStreamEndCapture(w->stream, &graph);
  JbWorld::GraphEntry e;
  e.k0 = k0; e.k1 = k1; e.k2 = k2; e.exec = nullptr; e.launches = w->launches - before;
  w->launches = before;  // counted when the graph is launched
  if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (ce != cudaSuccess) { jb_set_error(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce)); return JB_ECUDA; }
  CK(cudaGraphInstantiate(&e.exec, graph, 0));
  CK(cudaGraphDestroy(graph));
  CK(cudaGraphUpload(e.exec, w->stream));
  w->graphs.push_back(e);
  *index = w->graphs.size() - 1;
  return JB_OK;
}

template <typename F>
static int run_phase(JbWorld* w, int phase, const void* p0, const void* p1, F&& enqueue) {
  if (plain_launch(w)) return enqueue();
  size_t i = 0;
  TRY(ensure_graph(w, phase, p0, p1, enqueue, &i));
  CK(cudaGraphLaunch(w->graphs[i].exec, w->stream));
  w->launches += w->graphs[i].launches;
  return JB_OK;
}

// A step that failed while it was being enqueued may or may not have run k_step_begin (a failed graph
// capture runs nothing): put the device's count of begun steps, which selects the parameter ring slot, back
// in line with the host's.
static int resync_step_count(JbWorld* w, int rc) {
  cudaStreamSynchronize(w->stream);
  const uint32_t seq = (uint32_t)w->step_seq;
  cudaMemcpy(w->d_step_seq, &seq, 4, cudaMemcpyHostToDevice);
  return rc;
}

static int ensure_event_buffers(JbWorld* w, uint32_t flags) {
  if ((flags & JB_STEP_RECORD_EVENTS) && !w->ev_group) {
    TRY(dev_alloc(&w->ev_infected, w->newinf_cap));
    TRY(dev_alloc(&w->ev_infector, w->newinf_cap));
    TRY(dev_alloc(&w->ev_group, w->newinf_cap));
    TRY(dev_alloc(&w->ev_var, w->newinf_cap));
    drop_graphs(w);
  }
  if (!(flags & (JB_STEP_RECORD_EVENTS | JB_STEP_RECORD_HEALTH)) || w->hev_person) return JB_OK;
  // health events: at most four per infected person and step (symptoms, admission/transfer, recovery or death)
  w->hev_cap = (uint32_t)std::min<uint64_t>(4ull * w->N, 0x7FFFFFFFull);
  TRY(dev_alloc(&w->hev_person, w->hev_cap));
  TRY(dev_alloc(&w->hev_kind, w->hev_cap));
  TRY(dev_alloc(&w->hev_group, w->hev_cap));
  TRY(ensure_res_group(w));
  drop_graphs(w);
  return JB_OK;
}

extern "C" int jb_step_begin(JbWorld* w, const JbStepParams* p, void* d_send) {
  if (!w || !p) { jb_set_error("null argument"); return JB_EINVAL; }
  REQUIRE(!w->in_step, "jb_step_begin: previous step not finished");
  REQUIRE(p->beta != nullptr, "JbStepParams.beta is null");
  REQUIRE(p->delta_time > 0.0, "delta_time must be positive");
  CK(cudaSetDevice(w->device));
  cudaStream_t s = w->stream;
  w->cur = *p;
  TRY(timing_begin(w, s));
  if ((p->flags & JB_STEP_RECORD_LAMBDA) && !w->lambda) TRY(dev_alloc(&w->lambda, w->NP));
  TRY(ensure_event_buffers(w, p->flags));
  // per-step parameter block: written into a pinned ring slot, pulled by k_step_begin (the h2d payload of every step)
  const int slot = (int)(w->step_seq++ & 3);
  CK(cudaEventSynchronize(w->ev_step_buf[slot]));  // the copy that last used this slot has finished
  StepDev* h = reinterpret_cast<StepDev*>(w->h_step[slot]);
  h->now = p->now; h->dt = p->delta_time; h->t_now = p->now + p->delta_time; h->seed = p->seed;
  h->step = p->step; h->activity_mask = p->activity_mask; h->flags = p->flags; h->n_inject = (uint32_t)w->n_inject; h->leisure_table = p->leisure_table; h->epoch = (uint32_t)w->step_seq; h->tpar = w->trans_par;
  memcpy(w->h_step[slot] + sizeof(StepDev), p->beta, sizeof(double) * (size_t)w->n_beta * w->R);
  if (int rc = run_phase(w, 0, d_send, nullptr, [&]() { return enqueue_begin(w, d_send); })) return resync_step_count(w, rc);
  CK(cudaEventRecord(w->ev_step_buf[slot], s));  // k_step_begin has read the slot once this fires
  w->in_step = true;
  return JB_OK;
}

extern "C" int jb_step_interact(JbWorld* w, const void* d_recv, void* d_ret_send) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  REQUIRE(w->in_step, "jb_step_interact without jb_step_begin");
  CK(cudaSetDevice(w->device));
  return run_phase(w, 1, d_recv, d_ret_send, [&]() { return enqueue_interact(w, d_recv, d_ret_send); });
}

extern "C" int jb_step_finish(JbWorld* w, const void* d_ret_recv, JbStepStats* out) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  REQUIRE(w->in_step, "jb_step_finish without jb_step_begin");
  CK(cudaSetDevice(w->device));
  cudaStream_t s = w->stream;
  TRY(run_phase(w, 2, d_ret_recv, nullptr, [&]() { return enqueue_finish(w, d_ret_recv); }));
  TRY(timing_end(w, s));
  w->in_step = false;
  w->trans_par ^= 1u;  // the health update has filled the other transmission buffer
  if (out) {
    CK(cudaStreamSynchronize(s));
    const uint32_t* c = w->h_counters;
    memset(out, 0, sizeof(*out));
    out->n_new_infections = c[CNT_NEWINF];
    out->n_infected = c[CNT_INFECTED];
    out->n_dead_total = c[CNT_DEAD_TOTAL];
    out->n_recovered_step = c[CNT_RECOVERED_STEP];
    out->n_deaths_step = c[CNT_DEATHS_STEP];
    out->n_hospitalised = c[CNT_HOSP];
    out->n_icu = c[CNT_ICU];
    out->n_draws = c[CNT_DRAWS];
    out->n_no_activity = c[CNT_NOACT];
    out->n_slots_used = c[CNT_SLOTS_HW];
    out->n_halo_infected = c[CNT_HALO_INF];
    for (int a = 0; a < JB_N_ACT; ++a) out->n_by_activity[a] = c[CNT_ACT0 + a];
    if (c[CNT_NOACT]) {
      jb_set_error("Attention! Some people do not have an activity in this timestep.");
      return JB_ENOACT;
    }
    if (c[CNT_OVERFLOW]) {
      jb_set_error("a device buffer capacity was exceeded (dynamic members / new infections / slots)");
      return JB_ECAP;
    }
  }
  return JB_OK;
}

// Single-GPU step: the three phases are captured as ONE graph.
extern "C" int jb_step(JbWorld* w, const JbStepParams* p, JbStepStats* out) {
  if (w && p && halo_step_for(w, p->activity_mask) && !w->p2p_on) {  // host-issued collectives run between the phases
    TRY(jb_step_begin(w, p, nullptr));
    TRY(jb_step_interact(w, nullptr, nullptr));
    return jb_step_finish(w, nullptr, out);
  }
  // same prologue as jb_step_begin, then one graph for begin + interact + finish
  if (!w || !p) { jb_set_error("null argument"); return JB_EINVAL; }
  REQUIRE(!w->in_step, "jb_step: previous step not finished");
  REQUIRE(p->beta != nullptr, "JbStepParams.beta is null");
  REQUIRE(p->delta_time > 0.0, "delta_time must be positive");
  CK(cudaSetDevice(w->device));
  cudaStream_t s = w->stream;
  w->cur = *p;
  if ((p->flags & JB_STEP_RECORD_LAMBDA) && !w->lambda) TRY(dev_alloc(&w->lambda, w->NP));
  TRY(ensure_event_buffers(w, p->flags));
  auto whole_step = [&]() {
    TRY(enqueue_begin(w, nullptr));
    TRY(enqueue_interact(w, nullptr, nullptr));
    return enqueue_finish(w, nullptr);
  };
  const bool plain = plain_launch(w);
  size_t gi = 0;
  if (!plain) TRY(ensure_graph(w, 3, nullptr, nullptr, whole_step, &gi));  // first use of this key: captured before the bracket opens
  TRY(timing_begin(w, s));
  const int slot = (int)(w->step_seq++ & 3);
  CK(cudaEventSynchronize(w->ev_step_buf[slot]));
  StepDev* h = reinterpret_cast<StepDev*>(w->h_step[slot]);
  h->now = p->now; h->dt = p->delta_time; h->t_now = p->now + p->delta_time; h->seed = p->seed;
  h->step = p->step; h->activity_mask = p->activity_mask; h->flags = p->flags; h->n_inject = (uint32_t)w->n_inject; h->leisure_table = p->leisure_table; h->epoch = (uint32_t)w->step_seq; h->tpar = w->trans_par;
  memcpy(w->h_step[slot] + sizeof(StepDev), p->beta, sizeof(double) * (size_t)w->n_beta * w->R);
  if (plain) {
    if (int rc = whole_step()) return resync_step_count(w, rc);
  } else {
    CK(cudaGraphLaunch(w->graphs[gi].exec, s));
    w->launches += w->graphs[gi].launches;
  }
  CK(cudaEventRecord(w->ev_step_buf[slot], s));
  w->in_step = true;  // jb_step_finish_tail expects it
  TRY(timing_end(w, s));
  w->in_step = false;
  w->trans_par ^= 1u;  // the health update has filled the other transmission buffer
  if (!out) return JB_OK;
  CK(cudaStreamSynchronize(s));
  {
    const uint32_t* c = w->h_counters;
    memset(out, 0, sizeof(*out));
    out->n_new_infections = c[CNT_NEWINF];
    out->n_infected = c[CNT_INFECTED];
    out->n_dead_total = c[CNT_DEAD_TOTAL];
    out->n_recovered_step = c[CNT_RECOVERED_STEP];
    out->n_deaths_step = c[CNT_DEATHS_STEP];
    out->n_hospitalised = c[CNT_HOSP];
    out->n_icu = c[CNT_ICU];
    out->n_draws = c[CNT_DRAWS];
    out->n_no_activity = c[CNT_NOACT];
    out->n_slots_used = c[CNT_SLOTS_HW];
    out->n_halo_infected = c[CNT_HALO_INF];
    for (int a = 0; a < JB_N_ACT; ++a) out->n_by_activity[a] = c[CNT_ACT0 + a];
    if (c[CNT_NOACT]) {
      jb_set_error("Attention! Some people do not have an activity in this timestep.");
      return JB_ENOACT;
    }
    if (c[CNT_OVERFLOW]) {
      jb_set_error("a device buffer capacity was exceeded (dynamic members / new infections / slots)");
      return JB_ECAP;
    }
  }
  return JB_OK;
}

// Capture and instantiate the graph of a step with these parameters without running it, so that no later
// jb_step pays for a capture (bench / production set-up walks the schedule once).  Steps that exchange halo
// persons through host-issued collectives are captured phase by phase on first use instead.
extern "C" int jb_step_prepare(JbWorld* w, const JbStepParams* p) {
  if (!w || !p) { jb_set_error("null argument"); return JB_EINVAL; }
  REQUIRE(!w->in_step, "jb_step_prepare inside a step");
  CK(cudaSetDevice(w->device));
  const JbStepParams saved = w->cur;
  w->cur = *p;
  int rc = JB_OK;
  if (!plain_launch(w) && !(halo_step_for(w, p->activity_mask) && !w->p2p_on)) {
    rc = ensure_event_buffers(w, p->flags);
    size_t gi = 0;
    if (!rc)
      rc = ensure_graph(w, 3, nullptr, nullptr, [&]() {
        TRY(enqueue_begin(w, nullptr));
        TRY(enqueue_interact(w, nullptr, nullptr));
        return enqueue_finish(w, nullptr);
      }, &gi);
  }
  w->cur = saved;
  return rc;
}

// Evict the L2 between timed steps: a memset larger than the cache on the library's stream, in
// front of (= outside) the next step's event bracket, without a host synchronisation.
extern "C" int jb_flush_l2(JbWorld* w, size_t bytes) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  if (bytes > w->flush_bytes) {
    CK(cudaStreamSynchronize(w->stream));
    if (w->flush_buf) cudaFree(w->flush_buf);
    CK(cudaMalloc((void**)&w->flush_buf, bytes));
    w->flush_bytes = bytes;
  }
  CK(cudaMemsetAsync(w->flush_buf, 0, bytes, w->stream));
  return JB_OK;
}

// Human-readable summary of the device layout (one line per supergroup): which kernels process it and how
// its members' bytes are read.  Diagnostic only.
extern "C" int jb_world_describe(JbWorld* w, char* buf, size_t cap) {
  if (!w || !buf || !cap) { jb_set_error("null argument"); return JB_EINVAL; }
  std::string o;
  char line[512];
  snprintf(line, sizeof line, "people %u (internal %u, halo %u), groups %u, registrations %u, mirror slots %u\n", w->N_ext, w->N, w->H, w->G, w->M, w->mir_slots);
  o += line;
  static const char* modes[] = {"per-group kernels (gather)", "gather tiles", "stream tiles (slot == person)", "mirror (slot-ordered state)"};
  for (size_t k = 0; k < w->sgs.size(); ++k) {
    const SgTiles& T = w->tiles[k];
    snprintf(line, sizeof line, "supergroup %zu: groups %u, activity %d, dynamic %d, %s, tiles %u%s, span items %u, warp-list %u, cta-list %u, mean tiled group %.2f\n", k,
             w->sgs[k].n_groups, w->sgs[k].activity, w->sgs[k].has_dynamic, modes[T.mode], T.n_tiles, T.has_small ? "" : " (none valid)", T.n_span, T.n_w, T.n_c, T.mean_len);
    o += line;
  }
  snprintf(buf, cap, "%s", o.c_str());
  return JB_OK;
}

extern "C" int jb_sync(JbWorld* w) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  return JB_OK;
}

extern "C" void* jb_stream_handle(JbWorld* w) { return w ? (void*)w->stream : nullptr; }

// ---- readback ----------------------------------------------------------------------------
extern "C" int jb_fetch_new_infections(JbWorld* w, uint32_t* member, uint32_t* variant, size_t cap, size_t* n) {
  if (!w || !n) { jb_set_error("null argument"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  uint32_t cnt = 0;
  CK(cudaMemcpy(&cnt, w->counters + CNT_NEWINF, 4, cudaMemcpyDeviceToHost));
  cnt = std::min(cnt, w->newinf_cap);
  *n = cnt;
  REQUIRE(cnt <= cap, "jb_fetch_new_infections: caller buffer too small");
  if (cnt == 0) return JB_OK;
  std::vector<uint32_t> m(cnt);
  std::vector<uint8_t> v(cnt);
  CK(cudaMemcpy(m.data(), w->newinf_member, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(v.data(), w->newinf_var, cnt, cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < cnt; ++i)  // back to the caller's numbering (halo persons follow the residents)
    m[i] = m[i] < w->N ? w->h_int2ext[m[i]] : w->N_ext + (m[i] - w->N);
  // the device appends in arbitrary order: return sorted by member index
  std::vector<uint32_t> order(cnt);
  for (uint32_t i = 0; i < cnt; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return m[x] < m[y]; });
  for (uint32_t i = 0; i < cnt; ++i) {
    member[i] = m[order[i]];
    if (variant) variant[i] = v[order[i]];
  }
  return JB_OK;
}

extern "C" int jb_fetch_infection_events(JbWorld* w, uint32_t* infected, uint32_t* infector, uint32_t* group,
                                         uint32_t* variant, size_t cap, size_t* n) {
  if (!w || !n) { jb_set_error("null argument"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  *n = 0;
  if (!w->ev_group || !(w->cur.flags & JB_STEP_RECORD_EVENTS)) return JB_OK;  // the last step did not record events
  uint32_t cnt = 0;
  CK(cudaMemcpy(&cnt, w->counters + CNT_NEWINF, 4, cudaMemcpyDeviceToHost));
  cnt = std::min(cnt, w->newinf_cap);
  REQUIRE(cnt <= cap, "jb_fetch_infection_events: caller buffers too small");
  *n = cnt;
  if (!cnt) return JB_OK;
  std::vector<uint32_t> a(cnt), b(cnt), g(cnt);
  std::vector<uint8_t> v(cnt);
  CK(cudaMemcpy(a.data(), w->ev_infected, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), w->ev_infector, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(g.data(), w->ev_group, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(v.data(), w->ev_var, cnt, cudaMemcpyDeviceToHost));
  auto ext = [&](uint32_t p) { return p == 0xFFFFFFFFu ? p : p < w->N ? w->h_int2ext[p] : w->N_ext + (p - w->N); };
  std::vector<uint32_t> order(cnt);
  for (uint32_t i = 0; i < cnt; ++i) { order[i] = i; a[i] = ext(a[i]); b[i] = ext(b[i]); }
  std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return a[x] < a[y]; });
  for (uint32_t i = 0; i < cnt; ++i) {
    const uint32_t k = order[i];
    if (infected) infected[i] = a[k];
    if (infector) infector[i] = b[k];
    if (group) group[i] = g[k];
    if (variant) variant[i] = v[k];
  }
  return JB_OK;
}

extern "C" int jb_fetch_health_events(JbWorld* w, uint32_t* person, uint32_t* kind, int32_t* group, size_t cap, size_t* n) {
  if (!w || !n) { jb_set_error("null argument"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  *n = 0;
  if (!w->hev_person || !(w->cur.flags & (JB_STEP_RECORD_EVENTS | JB_STEP_RECORD_HEALTH))) return JB_OK;
  uint32_t cnt = 0;
  CK(cudaMemcpy(&cnt, w->counters + CNT_HEV, 4, cudaMemcpyDeviceToHost));
  cnt = std::min(cnt, w->hev_cap);
  REQUIRE(cnt <= cap, "jb_fetch_health_events: caller buffers too small");
  *n = cnt;
  if (!cnt) return JB_OK;
  std::vector<uint32_t> a(cnt), k(cnt), order(cnt);
  std::vector<int32_t> g(cnt);
  CK(cudaMemcpy(a.data(), w->hev_person, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(k.data(), w->hev_kind, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(g.data(), w->hev_group, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < cnt; ++i) { order[i] = i; a[i] = w->h_int2ext[a[i]]; }
  std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return a[x] != a[y] ? a[x] < a[y] : k[x] < k[y]; });
  for (uint32_t i = 0; i < cnt; ++i) {
    const uint32_t j = order[i];
    if (person) person[i] = a[j];
    if (kind) kind[i] = k[j];
    if (group) group[i] = g[j];
  }
  return JB_OK;
}

template <typename T>
static int fetch_permuted(JbWorld* w, const T* dev, T* out) {
  std::vector<T> tmp(w->N);
  CK(cudaMemcpy(tmp.data(), dev, (size_t)w->N * sizeof(T), cudaMemcpyDeviceToHost));
  for (uint32_t p = 0; p < w->N_ext; ++p) out[p] = tmp[w->h_ext2int[p]];
  return JB_OK;
}

extern "C" int jb_fetch_person_state(JbWorld* w, uint8_t* pstate, int8_t* tag, double* trans_prob,
                                     double* susceptibility, int32_t* medical) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  if (pstate) {  // ABI byte = activity of the last step (phot) | persistent flags (bits 3-6)
    std::vector<uint8_t> hot(w->N_ext);
    TRY(fetch_permuted(w, w->pstate, pstate));
    TRY(fetch_permuted(w, w->phot, hot.data()));
    for (uint32_t p = 0; p < w->N_ext; ++p) pstate[p] = (uint8_t)((pstate[p] & 0x78u) | (hot[p] & JB_PS_ACT_MASK));
  }
  if (tag) TRY(fetch_permuted(w, w->ptag, tag));
  if (trans_prob) {  // the buffer the next step reads; entries of people without an infection are stale by design
    std::vector<uint8_t> st(w->N_ext);
    TRY(fetch_permuted(w, w->trans[w->trans_par], trans_prob));
    TRY(fetch_permuted(w, w->pstate, st.data()));
    for (uint32_t p = 0; p < w->N_ext; ++p)
      if (!(st[p] & JB_PS_INFECTED)) trans_prob[p] = 0.0;
  }
  if (susceptibility)
    for (uint32_t v = 0; v < w->V; ++v) TRY(fetch_permuted(w, w->susc + (size_t)v * w->NP, susceptibility + (size_t)v * w->N_ext));
  if (medical) TRY(fetch_permuted(w, w->pmed, medical));
  return JB_OK;
}

extern "C" int jb_fetch_lambda(JbWorld* w, double* out) {
  if (!w || !out) { jb_set_error("null argument"); return JB_EINVAL; }
  REQUIRE(w->lambda != nullptr, "no step ran with JB_STEP_RECORD_LAMBDA");
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  std::vector<double> tmp(w->NP);
  CK(cudaMemcpy(tmp.data(), w->lambda, (size_t)w->NP * 8, cudaMemcpyDeviceToHost));
  for (uint32_t p = 0; p < w->N_ext; ++p) out[p] = tmp[w->h_ext2int[p]];
  for (uint32_t j = 0; j < w->H; ++j) out[w->N_ext + j] = tmp[w->N + j];
  return JB_OK;
}

extern "C" int jb_fetch_infections(JbWorld* w, JbInfectionRow* rows, size_t cap, size_t* n) {
  if (!w || !n) { jb_set_error("null argument"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  uint32_t hw = 0;
  CK(cudaMemcpy(&hw, w->counters + CNT_SLOTS_HW, 4, cudaMemcpyDeviceToHost));
  hw = std::min<uint32_t>(hw, w->C);
  const size_t C = w->C;
  std::vector<int32_t> person(hw);
  std::vector<double> start(hw), par((size_t)hw * JB_N_TPAR), tt((size_t)hw * JB_MAX_STAGES), trans(w->N);
  std::vector<uint8_t> stage(hw), nst(hw), var(hw);
  std::vector<int8_t> tag(hw), maxtag(hw), ttag((size_t)hw * JB_MAX_STAGES);
  if (hw) {
    CK(cudaMemcpy(person.data(), w->inf_person, (size_t)hw * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(start.data(), w->inf_start, (size_t)hw * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(stage.data(), w->inf_stage, hw, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(nst.data(), w->inf_nst, hw, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(var.data(), w->inf_var, hw, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(tag.data(), w->inf_tag, hw, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(maxtag.data(), w->inf_maxtag, hw, cudaMemcpyDeviceToHost));
    for (int q = 0; q < JB_N_TPAR; ++q)
      CK(cudaMemcpy(par.data() + (size_t)q * hw, w->inf_par + (size_t)q * C, (size_t)hw * 8, cudaMemcpyDeviceToHost));
    for (int q = 0; q < JB_MAX_STAGES; ++q) {
      CK(cudaMemcpy(tt.data() + (size_t)q * hw, w->traj_time + (size_t)q * C, (size_t)hw * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(ttag.data() + (size_t)q * hw, w->traj_tag + (size_t)q * C, hw, cudaMemcpyDeviceToHost));
    }
    CK(cudaMemcpy(trans.data(), w->trans[w->trans_par], (size_t)w->N * 8, cudaMemcpyDeviceToHost));
  }
  size_t k = 0;
  for (uint32_t s = 0; s < hw; ++s) {
    if (person[s] < 0) continue;
    if (k < cap && rows) {
      JbInfectionRow& r = rows[k];
      memset(&r, 0, sizeof(r));
      r.person = w->h_int2ext[person[s]]; r.variant = var[s]; r.start_time = start[s];
      for (int q = 0; q < JB_N_TPAR; ++q) r.tpar[q] = par[(size_t)q * hw + s];
      if (w->dis_host.trans_type == JB_TRANS_GAMMA) r.tpar[4] = 0.0;  // (device-internal: the profile's constant factor)
      r.probability = trans[person[s]];
      r.n_stages = nst[s]; r.stage = stage[s]; r.max_tag = maxtag[s]; r.tag = tag[s];
      for (int q = 0; q < JB_MAX_STAGES; ++q) { r.traj_time[q] = tt[(size_t)q * hw + s]; r.traj_tag[q] = ttag[(size_t)q * hw + s]; }
    }
    ++k;
  }
  *n = k;
  REQUIRE(!rows || k <= cap, "jb_fetch_infections: caller buffer too small");
  if (rows) std::sort(rows, rows + k, [](const JbInfectionRow& a, const JbInfectionRow& b) { return a.person < b.person; });
  return JB_OK;
}

extern "C" int jb_fetch_summary(JbWorld* w, uint32_t* counts) {
  if (!w || !counts) { jb_set_error("null argument"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));  // the d2h copy of the last step has landed in h_counters
  memcpy(counts, w->h_counters + CNT_N, sizeof(uint32_t) * (size_t)w->R * JB_N_SUMMARY);
  return JB_OK;
}

// on = 0: run the supergroup kernels one after the other on the main stream (isolated per-kernel
// timings for the roofline); on = 1 (default): concurrently on side streams.
extern "C" int jb_set_concurrency(JbWorld* w, int on) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  CK(cudaStreamSynchronize(w->stream));
  w->serial = on == 0;
  return JB_OK;
}

extern "C" int jb_timing_enable(JbWorld* w, int on) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  TRY(harvest_timing(w));
  w->timing = on;  // 1: whole-step events only (graphs stay on); 2: + per-kernel events (plain launches)
  w->ev_detail = on >= 2;
  return JB_OK;
}

extern "C" int jb_timing_read(JbWorld* w, double* interaction_ms, double* step_ms, uint64_t* n_steps,
                              uint64_t* n_launches, int reset) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  TRY(harvest_timing(w));
  if (interaction_ms) *interaction_ms = w->acc_int_ms;
  if (step_ms) *step_ms = w->acc_step_ms;
  if (n_steps) *n_steps = w->acc_steps;
  if (n_launches) *n_launches = w->launches;
  if (reset) {
    w->acc_int_ms = w->acc_step_ms = 0; w->acc_steps = 0; w->launches = 0;
    for (int k = 0; k < JB_MAX_SUPERGROUPS; ++k) { w->acc_sg_ms[k] = 0; w->acc_sg_launches[k] = 0; }
  }
  return JB_OK;
}

// Per-supergroup interaction kernel time (ms, launches) accumulated since the last reset.
extern "C" int jb_timing_read_supergroups(JbWorld* w, double* ms, uint64_t* launches, int cap) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  CK(cudaSetDevice(w->device));
  TRY(harvest_timing(w));
  for (int k = 0; k < cap && k < (int)w->sgs.size(); ++k) { ms[k] = w->acc_sg_ms[k]; launches[k] = w->acc_sg_launches[k]; }
  return JB_OK;
}

extern "C" int jb_last_step_bytes(JbWorld* w, uint64_t* interaction_bytes, uint64_t* step_bytes) {
  if (!w) { jb_set_error("null world"); return JB_EINVAL; }
  if (interaction_bytes) *interaction_bytes = w->last_int_bytes;
  if (step_bytes) *step_bytes = w->last_step_bytes;
  return JB_OK;
}
