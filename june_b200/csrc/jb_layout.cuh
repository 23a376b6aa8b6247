// World-create-time layout: internal person numbering and the 32-slot tile decomposition.
//
// The reference keeps people in Python lists per subgroup; the device wants the dominant access
// (every resident's household, every step) to be a coalesced stream.  So residents are renumbered
// by the padded slot order of the streamed residence supergroup: person index == slot index, no
// member indirection for households.  Padding slots are permanently "dead" pseudo-persons.
// Everything crossing the C ABI is translated back to the caller's numbering.
#pragma once
#include <algorithm>
#include <vector>

#include "jb_internal.cuh"


struct HostTiles {
  int mode = 0;  // 0 warp only, 1 gather tiles, 2 stream tiles (slot == person), 3 mirror tiles (slot-ordered state copy)
  std::vector<uint32_t> tile_group0, tp_member, big;
  std::vector<uint32_t> big_slot0;  // stream / mirror: first slot of every group in `big`
  std::vector<uint8_t> tp_info;
  // span layout (k_span): work items {first slot, kind | count, first entry of span_group, 0} and the groups in slot order
  bool span = false;
  std::vector<uint4> span_items;
  std::vector<uint32_t> span_group;
};

// Slot layout of a mirrored supergroup with one subgroup per group, built for k_span: every warp-sized work item is
// full.  Groups of more than 512 slots first (longest first, whole 32-slot tiles each), then the size classes
// 257-512 / 129-256 / 65-128 / 33-64 (every group padded to 512 / 256 / 128 / 64 slots, so that 1 / 2 / 4 / 8 groups
// fill the 512 slots a warp loads at once), then the groups of <= 32 slots packed into tiles that never split a group.
// Groups above `big_max` slots are laid out like the others but left to the CTA-per-group kernel (no work item).
static void jb_pack_span(const JbWorldDesc* d, const JbSupergroup& sg, uint32_t big_max, HostTiles& T) {
  T.span = true;
  auto gsize = [&](uint32_t g) { return d->grp_offset[g + 1] - d->grp_offset[g]; };
  auto pad_to = [&](size_t mult) {
    while (T.tp_member.size() % mult) { T.tp_member.push_back(0xFFFFFFFFu); T.tp_info.push_back(0); }
  };
  auto put_group = [&](uint32_t g) {  // members in registration order; not VALID: no tile logic applies
    for (uint32_t m = d->grp_offset[g]; m < d->grp_offset[g + 1]; ++m) {
      T.tp_member.push_back(d->reg_member[m]);
      T.tp_info.push_back((uint8_t)JB_INFO_ACT(d->reg_info[m]));
    }
  };
  std::vector<uint32_t> cls[5], small;  // cls[0]: > 512, cls[1]: 257-512 (32 lanes), cls[2]: 16 lanes, cls[3]: 8, cls[4]: 4
  for (uint32_t g = sg.first_group; g < sg.first_group + sg.n_groups; ++g) {
    const uint32_t sz = gsize(g);
    if (sz == 0) continue;
    if (sz <= 32) small.push_back(g);
    else cls[sz > 512 ? 0 : sz > 256 ? 1 : sz > 128 ? 2 : sz > 64 ? 3 : 4].push_back(g);
  }
  std::stable_sort(cls[0].begin(), cls[0].end(), [&](uint32_t x, uint32_t y) { return gsize(x) > gsize(y); });
  for (uint32_t g : cls[0]) {
    pad_to(32);
    T.big.push_back(g); T.big_slot0.push_back((uint32_t)T.tp_member.size());
    if (gsize(g) <= big_max) {
      T.span_items.push_back(make_uint4((uint32_t)T.tp_member.size(), JB_SPAN_KIND_BIG << 28, (uint32_t)T.span_group.size(), 0u));
      T.span_group.push_back(g);
    }
    put_group(g);
  }
  for (int c = 1; c <= 4; ++c) {
    const uint32_t lanes = 64u >> c, per_span = 32u / lanes;  // 32, 16, 8, 4 lanes per group
    for (size_t i = 0; i < cls[c].size(); i += per_span) {
      pad_to(512);
      const uint32_t cnt = (uint32_t)std::min<size_t>(per_span, cls[c].size() - i);
      T.span_items.push_back(make_uint4((uint32_t)T.tp_member.size(), (JB_SPAN_KIND_CLASS << 28) | (lanes << 8) | cnt, (uint32_t)T.span_group.size(), 0u));
      for (uint32_t j = 0; j < cnt; ++j) {
        const uint32_t g = cls[c][i + j];
        pad_to((size_t)lanes * 16);
        T.big.push_back(g); T.big_slot0.push_back((uint32_t)T.tp_member.size());
        T.span_group.push_back(g);
        put_group(g);
      }
    }
  }
  pad_to(512);
  // small groups: tiles that never split a group, 16 tiles per work item
  const size_t small0 = T.tp_member.size();
  T.tile_group0.assign(small0 / 32, 0u);
  uint32_t fill = 32;
  for (uint32_t g : small) {
    const uint32_t sz = gsize(g);
    if (fill + sz > 32) { pad_to(32); T.tile_group0.push_back((uint32_t)T.span_group.size()); fill = 0; }
    T.span_group.push_back(g);
    for (uint32_t m = d->grp_offset[g]; m < d->grp_offset[g + 1]; ++m) {
      T.tp_member.push_back(d->reg_member[m]);
      T.tp_info.push_back((uint8_t)(JB_INFO_ACT(d->reg_info[m]) | JB_TI_VALID | (m == d->grp_offset[g] ? JB_TI_FIRST : 0u)));
    }
    fill += sz;
  }
  pad_to(32);
  const uint32_t n_small_tiles = (uint32_t)((T.tp_member.size() - small0) / 32);
  for (uint32_t t = 0; t < n_small_tiles; t += 16)
    T.span_items.push_back(make_uint4((uint32_t)small0 + t * 32, (JB_SPAN_KIND_SMALL << 28) | std::min(16u, n_small_tiles - t), 0u, 0u));
  pad_to(512);  // the last work item may load a whole span
  T.tile_group0.resize(T.tp_member.size() / 32, 0u);
}

struct HostLayout {
  uint32_t N_int = 0;
  std::vector<uint32_t> ext2int, int2ext;
  std::vector<uint8_t> age, sex, ch, phas;
  std::vector<uint32_t> psa, reg_member;
  std::vector<HostTiles> tiles;
  // Mirror: slot-ordered copy of the person state for the supergroups of the primary activity whose members
  // are registered once (companies, schools).  One slot space over all mirrored supergroups:
  // supergroup k owns [mir_base[k], mir_base[k] + tiles[k].tp_member.size()).
  std::vector<uint32_t> mir_base;   // per supergroup (0xFFFFFFFF = not mirrored)
  std::vector<uint32_t> pmslot;     // [N_int + H] slot of the person's mirrored registration, or 0xFFFFFFFF
  uint32_t mir_slots = 0;
};

// Groups of more than `tile_max` slots are left to the per-group kernels (a tile lane group walks
// its group in rounds, which only pays for short groups).
static void jb_pack_tiles(const JbWorldDesc* d, const JbSupergroup& sg, bool stream, uint32_t tile_max, HostTiles& T) {
  auto pad = [&]() {
    while (T.tp_member.size() % 32) { T.tp_member.push_back(0xFFFFFFFFu); T.tp_info.push_back(0); }
  };
  uint32_t fill = 32;  // 32 = no open tile
  for (uint32_t g = sg.first_group; g < sg.first_group + sg.n_groups; ++g) {
    const uint32_t b = d->grp_offset[g], e = d->grp_offset[g + 1], sz = e - b;
    if (sz == 0) { pad(); fill = 32; continue; }  // keeps group indices consecutive inside a tile
    if (sz > tile_max) {
      T.big.push_back(g);
      pad();
      fill = 32;
      if (stream) {  // big groups still own a contiguous block of person indices (not tile-valid)
        T.big_slot0.push_back((uint32_t)T.tp_member.size());
        for (uint32_t m = b; m < e; ++m) {
          if ((m - b) % 32 == 0) T.tile_group0.push_back(g);
          T.tp_member.push_back(d->reg_member[m]);
          // not VALID: handled by the per-group kernels (the span kernel reads the subgroup bits)
          T.tp_info.push_back((uint8_t)(JB_INFO_ACT(d->reg_info[m]) | ((JB_INFO_LSG(d->reg_info[m]) & 3u) << 5)));
        }
        pad();
      }
      continue;
    }
    if (fill + sz > 32) { pad(); T.tile_group0.push_back(g); fill = 0; }
    for (uint32_t m = b; m < e; ++m) {
      T.tp_member.push_back(d->reg_member[m]);
      T.tp_info.push_back((uint8_t)(JB_INFO_ACT(d->reg_info[m]) | JB_TI_VALID | (m == b ? JB_TI_FIRST : 0u) |
                                    ((JB_INFO_LSG(d->reg_info[m]) & 3u) << 5)));
    }
    fill += sz;
  }
  pad();
}

// Returns an error string or nullptr.
static const char* jb_build_layout(const JbWorldDesc* d, const std::vector<JbSupergroup>& sgs, uint32_t tile_max_gather,
                                   bool mirror_on, bool span_on, uint32_t span_big_max, HostLayout& L) {
  const uint32_t N = d->n_people;
  L.tiles.assign(sgs.size(), HostTiles());
  L.mir_base.assign(sgs.size(), 0xFFFFFFFFu);
  // Mirrored supergroups: primary activity, static members only, every slot fed by the primary activity, every
  // resident member has the primary activity among its fixed ones, nobody registered twice over all of them
  // (the reference's Person.subgroups holds ONE primary_activity subgroup, person.py:20-27).  Presence of a
  // member can then be decided from its state byte and the step's activity mask alone.
  std::vector<uint8_t> mirrored(sgs.size(), 0);
  if (mirror_on) {
    std::vector<uint8_t> seen((size_t)N + d->n_halo, 0);
    for (size_t k = 0; k < sgs.size(); ++k) {
      const JbSupergroup& sg = sgs[k];
      if (sg.activity != JB_ACT_PRIMARY || sg.has_dynamic || sg.n_groups == 0) continue;
      const uint32_t m0 = d->grp_offset[sg.first_group], m1 = d->grp_offset[sg.first_group + sg.n_groups];
      bool ok = m1 > m0;
      for (uint32_t m = m0; ok && m < m1; ++m) {
        const uint32_t p = d->reg_member[m];
        if (JB_INFO_ACT(d->reg_info[m]) != JB_ACT_PRIMARY || p >= N + d->n_halo || seen[p]) ok = false;
        else if (p < N && !(d->has_activity[p] & (1u << JB_ACT_PRIMARY))) ok = false;
      }
      if (!ok) continue;
      for (uint32_t m = m0; m < m1; ++m) seen[d->reg_member[m]] = 1;
      mirrored[k] = 1;
    }
  }
  // the streamed supergroup: first tile-mode residence supergroup whose slots are distinct residents
  int ks = -1;
  for (size_t k = 0; k < sgs.size() && ks < 0; ++k) {
    const JbSupergroup& sg = sgs[k];
    if (sg.launch_mode != 0 || sg.activity != JB_ACT_RESIDENCE || sg.n_groups == 0) continue;
    std::vector<uint8_t> seen(N, 0);
    bool ok = true;
    for (uint32_t m = d->grp_offset[sg.first_group]; ok && m < d->grp_offset[sg.first_group + sg.n_groups]; ++m) {
      const uint32_t p = d->reg_member[m];
      if (p >= N || seen[p] || JB_INFO_ACT(d->reg_info[m]) != JB_ACT_RESIDENCE) ok = false; else seen[p] = 1;
    }
    if (ok) ks = (int)k;
  }
  for (size_t k = 0; k < sgs.size(); ++k) {
    const JbSupergroup& sg = sgs[k];
    if (sg.launch_mode != 0) {
      // per-group kernels only; a mirrored supergroup still gets slots: every group owns a contiguous block
      if (mirrored[k]) { L.tiles[k].mode = 3; jb_pack_tiles(d, sg, true, 0u, L.tiles[k]); }
      continue;
    }
    for (uint32_t m = d->grp_offset[sg.first_group]; m < d->grp_offset[sg.first_group + sg.n_groups]; ++m)
      if ((int)JB_INFO_LSG(d->reg_info[m]) >= sg.dim) return "registered member has a subgroup index >= contact matrix dimension";
    L.tiles[k].mode = ((int)k == ks) ? 2 : mirrored[k] ? 3 : 1;
    if (mirrored[k] && span_on && sg.kind == JB_KIND_MATRIX && sg.dim == 1) { jb_pack_span(d, sg, span_big_max, L.tiles[k]); continue; }
    jb_pack_tiles(d, sg, (int)k == ks || mirrored[k], ((int)k == ks || mirrored[k]) ? 32u : tile_max_gather, L.tiles[k]);
  }
  // internal numbering
  L.ext2int.assign(N, 0xFFFFFFFFu);
  uint32_t next = 0;
  if (ks >= 0) {
    const HostTiles& T = L.tiles[ks];
    for (uint32_t s = 0; s < T.tp_member.size(); ++s)
      if (T.tp_member[s] != 0xFFFFFFFFu) L.ext2int[T.tp_member[s]] = s;
    next = (uint32_t)T.tp_member.size();
  }
  for (uint32_t p = 0; p < N; ++p)
    if (L.ext2int[p] == 0xFFFFFFFFu) L.ext2int[p] = next++;
  L.N_int = next;
  L.int2ext.assign(L.N_int, 0xFFFFFFFFu);
  for (uint32_t p = 0; p < N; ++p) L.int2ext[L.ext2int[p]] = p;
  // remapped person attributes (padding: dead pseudo-persons without activities)
  L.age.assign(L.N_int, 0); L.sex.assign(L.N_int, 0); L.ch.assign(L.N_int, 0); L.phas.assign(L.N_int, 0);
  L.psa.assign(L.N_int, 0);
  for (uint32_t p = 0; p < N; ++p) {
    const uint32_t i = L.ext2int[p];
    L.age[i] = d->age[p]; L.sex[i] = d->sex[p]; L.ch[i] = d->care_home_resident[p];
    L.phas[i] = d->has_activity[p]; L.psa[i] = d->super_area[p];
  }
  auto remap = [&](uint32_t p) { return p < N ? L.ext2int[p] : L.N_int + (p - N); };
  L.reg_member.resize(d->n_members);
  for (uint32_t m = 0; m < d->n_members; ++m) L.reg_member[m] = remap(d->reg_member[m]);
  for (size_t k = 0; k < sgs.size(); ++k)
    if (L.tiles[k].mode == 1 || L.tiles[k].mode == 3)
      for (uint32_t& x : L.tiles[k].tp_member)
        if (x != 0xFFFFFFFFu) x = remap(x);
  // mirror slot space + person -> slot map
  L.pmslot.assign((size_t)L.N_int + d->n_halo, 0xFFFFFFFFu);
  for (size_t k = 0; k < sgs.size(); ++k) {
    if (L.tiles[k].mode != 3) continue;
    L.mir_base[k] = L.mir_slots;
    const HostTiles& T = L.tiles[k];
    for (uint32_t s = 0; s < T.tp_member.size(); ++s)
      if (T.tp_member[s] != 0xFFFFFFFFu) L.pmslot[T.tp_member[s]] = L.mir_slots + s;
    L.mir_slots += (uint32_t)T.tp_member.size();
  }
  return nullptr;
}
