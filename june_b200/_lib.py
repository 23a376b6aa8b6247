"""ctypes binding of ``libjune_b200.so`` (C ABI: ``include/june_b200.h``).

The product path has NO CPU fallback: if the shared library is missing, or no CUDA device is
usable, creating a device world raises.  The same structures are reused by the tests to drive
the CPU checker library under ``oracle/`` (same ABI, prefix ``jo_``) -- that library is never
loaded from inside this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional

import numpy as np

from .world import MAX_STAGES, N_TPAR

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libjune_b200.so")
CSRC = os.path.join(HERE, "csrc")
ABI_VERSION = 1

JB_OK, JB_EINVAL, JB_ECUDA, JB_ENOACT, JB_ECAP, JB_ENOGPU = 0, 1, 2, 3, 4, 5
STEP_SEVERE_STAY_HOME, STEP_HOSPITALISATION, STEP_INJECT_UNIFORMS = 0x1, 0x2, 0x4
STEP_NO_NEW_INFECTIONS, STEP_RECORD_LAMBDA, STEP_INJECT_LEISURE, STEP_INJECT_POLICY = 0x8, 0x10, 0x20, 0x40
STEP_RECORD_EVENTS, STEP_INJECT_COMMUTE, STEP_RECORD_HEALTH = 0x80, 0x100, 0x200
MAX_POLICIES, POLICY_DRAWS = 8, 3
POL_SEVERE_STAY_HOME, POL_QUARANTINE, POL_SHIELDING, POL_CLOSE_SCHOOLS, POL_CLOSE_UNIVERSITIES, POL_CLOSE_COMPANIES = 1, 2, 3, 4, 5, 6
POL_LIMIT_LONG_COMMUTE, POL_CLOSE_COMPANIES_TIERS = 7, 8


class SimulatorError(BaseException):
    """Same base class as the reference's ``june.exc.SimulatorError`` (``exc.py:13``)."""


class JbDist(C.Structure):
    _fields_ = [("type", C.c_int32), ("p", C.c_double * 4)]


class JbSupergroup(C.Structure):
    _fields_ = [("first_group", C.c_uint32), ("n_groups", C.c_uint32), ("kind", C.c_int32), ("dim", C.c_int32),
                ("cm_offset", C.c_int32), ("activity", C.c_int32), ("launch_mode", C.c_int32),
                ("has_dynamic", C.c_int32)]


class JbDisease(C.Structure):
    _fields_ = [("n_tags", C.c_int32), ("tag_recovered_mask", C.c_int32), ("tag_dead_mask", C.c_int32),
                ("tag_hospital_mask", C.c_int32), ("tag_icu_mask", C.c_int32), ("tag_severe_mask", C.c_int32),
                ("tag_exposed", C.c_int32), ("n_traj", C.c_int32),
                ("traj_max_tag", C.c_int32 * 16), ("traj_n_stages", C.c_int32 * 16),
                ("traj_stage_tag", (C.c_int32 * MAX_STAGES) * 16),
                ("traj_stage_time", (JbDist * MAX_STAGES) * 16),
                ("trans_type", C.c_int32), ("trans_par", JbDist * 6),
                ("health_index_cum", C.c_void_p), ("infection_id", C.c_uint32)]


class JbWorldDesc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32),
                ("n_people", C.c_uint32), ("n_halo", C.c_uint32), ("n_groups", C.c_uint32), ("n_members", C.c_uint32),
                ("n_supergroups", C.c_uint32), ("n_regions", C.c_uint32), ("n_super_areas", C.c_uint32),
                ("n_variants", C.c_uint32), ("n_beta", C.c_uint32), ("n_scale", C.c_uint32),
                ("n_cm_doubles", C.c_uint32), ("n_school_years", C.c_uint32),
                ("slot_capacity", C.c_uint32), ("newinf_capacity", C.c_uint32),
                ("person_id_base", C.c_uint64),
                ("age", C.c_void_p), ("sex", C.c_void_p), ("care_home_resident", C.c_void_p),
                ("has_activity", C.c_void_p), ("super_area", C.c_void_p),
                ("sa_hospital", C.c_void_p), ("sa_region", C.c_void_p),
                ("grp_offset", C.c_void_p), ("grp_info", C.c_void_p), ("grp_aux", C.c_void_p),
                ("reg_member", C.c_void_p), ("reg_info", C.c_void_p), ("school_years", C.c_void_p),
                ("supergroups", C.c_void_p), ("cm", C.c_void_p), ("beta_scale", C.c_void_p),
                ("disease", JbDisease)]


class JbStepParams(C.Structure):
    _fields_ = [("now", C.c_double), ("delta_time", C.c_double), ("step", C.c_uint32), ("activity_mask", C.c_uint32),
                ("seed", C.c_uint64), ("flags", C.c_uint32), ("leisure_table", C.c_uint32), ("beta", C.c_void_p)]


class JbIndividualPolicy(C.Structure):
    _fields_ = [("type", C.c_int32), ("flags", C.c_uint32), ("mask", C.c_uint32), ("_pad", C.c_uint32),
                ("p", C.c_double * 6)]


MAX_DOSES = 4


class JbVaccine(C.Structure):
    _fields_ = [("n_doses", C.c_int32), ("days_administered_to_effective", C.c_int32 * MAX_DOSES),
                ("days_effective_to_waning", C.c_int32 * MAX_DOSES), ("days_waning", C.c_int32 * MAX_DOSES),
                ("waning_factor", C.c_double), ("sterilisation", C.c_void_p), ("symptomatic", C.c_void_p)]


class JbVisitsDesc(C.Structure):
    _fields_ = [("activity", C.c_int32), ("household_supergroup", C.c_int32), ("care_home_supergroup", C.c_int32),
                ("care_home_subgroup", C.c_int32), ("household_off", C.c_void_p), ("household_target", C.c_void_p),
                ("care_home_off", C.c_void_p), ("care_home_target", C.c_void_p), ("p_household", C.c_double),
                ("p_care_home", C.c_double), ("household_visits_beta", C.c_int32)]


class JbVaccinationState(C.Structure):
    _fields_ = [("dose", C.c_void_p), ("vaccine", C.c_void_p), ("campaign", C.c_void_p), ("first_day", C.c_void_p),
                ("stage", C.c_void_p), ("prior_infection", C.c_void_p), ("prior_symptoms", C.c_void_p),
                ("effective_multiplier", C.c_void_p)]


class JbVaccinationCampaign(C.Structure):
    _fields_ = [("vaccine", C.c_int32), ("group_by", C.c_int32), ("lo", C.c_int32), ("hi", C.c_int32),
                ("n_doses", C.c_int32), ("dose_numbers", C.c_int32 * MAX_DOSES),
                ("days_to_next_dose", C.c_int32 * MAX_DOSES), ("start_day", C.c_int32), ("end_day", C.c_int32),
                ("last_dose_vaccines", C.c_uint32), ("coverage", C.c_double)]


class JbCommuteDesc(C.Structure):
    _fields_ = [("n_cities", C.c_uint32), ("n_stations", C.c_uint32), ("person_commute", C.c_void_p),
                ("city_station_off", C.c_void_p), ("city_station", C.c_void_p),
                ("station_transport_off", C.c_void_p), ("station_transport", C.c_void_p)]


class JbLeisureDesc(C.Structure):
    _fields_ = [("n_activities", C.c_uint32), ("n_areas", C.c_uint32), ("supergroup", C.c_void_p),
                ("subgroup", C.c_void_p), ("person_area", C.c_void_p), ("area_off", C.c_void_p),
                ("area_venue", C.c_void_p)]


class JbStepStats(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "n_new_infections", "n_infected", "n_dead_total", "n_recovered_step", "n_deaths_step", "n_hospitalised",
        "n_icu", "n_draws", "n_no_activity", "n_slots_used", "n_halo_infected")] + [("n_by_activity", C.c_uint32 * 5)]

    def as_dict(self):
        # (one unpack of the 16 words instead of a getattr per field: this runs in every end-to-end step)
        words = _STATS_STRUCT.unpack_from(self)
        d = dict(zip(_STATS_NAMES, words))
        d["n_by_activity"] = list(words[len(_STATS_NAMES):])
        return d


_STATS_NAMES = tuple(n for n, _ in JbStepStats._fields_ if n != "n_by_activity")
_STATS_STRUCT = __import__("struct").Struct("=%dI" % (len(_STATS_NAMES) + 5))
assert _STATS_STRUCT.size == C.sizeof(JbStepStats)


class JbInfectionRow(C.Structure):
    _fields_ = [("person", C.c_uint32), ("variant", C.c_uint32), ("start_time", C.c_double),
                ("tpar", C.c_double * N_TPAR), ("probability", C.c_double),
                ("n_stages", C.c_int32), ("stage", C.c_int32), ("max_tag", C.c_int32), ("tag", C.c_int32),
                ("traj_time", C.c_double * MAX_STAGES), ("traj_tag", C.c_int32 * MAX_STAGES)]


# every entry point include/june_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "last_error", "abi_version", "device_count", "world_create", "world_destroy", "set_person_state",
    "inject_uniforms", "infect", "upload_infections", "step", "step_begin", "step_interact", "step_finish",
    "set_halo", "halo_record_size", "halo_return_size", "sync", "fetch_new_infections", "fetch_person_state",
    "fetch_lambda", "fetch_infections", "fetch_summary", "timing_enable", "timing_read", "last_step_bytes",
    "set_concurrency", "set_leisure", "set_leisure_tables", "inject_leisure",
    "set_individual_policies", "set_person_attributes", "inject_policy_uniforms", "fetch_infection_events",
    "set_commute", "inject_commute",
    "set_vaccination", "vaccinate", "inject_vaccination_uniforms", "fetch_vaccination",
    "p2p_window", "p2p_connect", "flush_l2", "fetch_health_events", "set_medical", "set_halo_activities", "set_variants", "mutate", "set_person_work_regions",
    "step_prepare", "world_describe", "vaccination_state", "set_visits", "set_visit_probabilities",
]


def nvcc_command(out_path: str = LIB_PATH) -> List[str]:
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
            "-Xcompiler", "-fPIC", "-o", out_path, os.path.join(CSRC, "june_b200.cu")]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "june_b200.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    cmd = nvcc_command()
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


def bind(lib: C.CDLL, prefix: str) -> C.CDLL:
    """Declare argument / return types of all ABI functions on ``lib`` (``jb_`` or ``jo_``)."""
    def f(name):
        return getattr(lib, prefix + name)
    vp, u32p, u8p, dp = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_double)
    f("last_error").restype = C.c_char_p
    f("last_error").argtypes = []
    f("abi_version").restype = C.c_int
    f("device_count").restype = C.c_int
    f("world_create").argtypes = [C.POINTER(JbWorldDesc), C.c_int, C.POINTER(vp)]
    f("world_destroy").argtypes = [vp]
    f("world_destroy").restype = None
    f("set_person_state").argtypes = [vp, vp, vp]
    f("inject_uniforms").argtypes = [vp, vp, C.c_size_t]
    f("infect").argtypes = [vp, vp, vp, C.c_size_t, C.c_double, C.c_uint64, C.c_uint32]
    f("upload_infections").argtypes = [vp, C.POINTER(JbInfectionRow), C.c_size_t]
    f("step").argtypes = [vp, C.POINTER(JbStepParams), C.POINTER(JbStepStats)]
    f("step_prepare").argtypes = [vp, C.POINTER(JbStepParams)]
    f("world_describe").argtypes = [vp, C.c_char_p, C.c_size_t]
    f("step_begin").argtypes = [vp, C.POINTER(JbStepParams), vp]
    f("step_interact").argtypes = [vp, vp, vp]
    f("step_finish").argtypes = [vp, vp, C.POINTER(JbStepStats)]
    f("set_halo").argtypes = [vp, vp, C.c_size_t]
    f("halo_record_size").restype = C.c_size_t
    f("halo_return_size").restype = C.c_size_t
    f("sync").argtypes = [vp]
    f("fetch_new_infections").argtypes = [vp, vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    f("fetch_person_state").argtypes = [vp, vp, vp, vp, vp, vp]
    f("fetch_lambda").argtypes = [vp, vp]
    f("fetch_infections").argtypes = [vp, C.POINTER(JbInfectionRow), C.c_size_t, C.POINTER(C.c_size_t)]
    f("fetch_summary").argtypes = [vp, vp]
    f("timing_enable").argtypes = [vp, C.c_int]
    f("timing_read").argtypes = [vp, dp, dp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]
    f("last_step_bytes").argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    f("set_leisure").argtypes = [vp, C.POINTER(JbLeisureDesc)]
    f("set_leisure_tables").argtypes = [vp, C.c_uint32, vp, vp]
    f("inject_leisure").argtypes = [vp, vp]
    f("set_individual_policies").argtypes = [vp, C.POINTER(JbIndividualPolicy), C.c_uint32, vp]
    f("set_person_attributes").argtypes = [vp, vp]
    f("inject_policy_uniforms").argtypes = [vp, vp]
    f("set_vaccination").argtypes = [vp, C.POINTER(JbVaccine), C.c_uint32, C.POINTER(JbVaccinationCampaign), C.c_uint32]
    f("vaccinate").argtypes = [vp, C.c_int32, C.c_uint64]
    f("inject_vaccination_uniforms").argtypes = [vp, vp]
    f("fetch_vaccination").argtypes = [vp, vp, vp, vp]
    f("vaccination_state").argtypes = [vp, C.c_int, C.POINTER(JbVaccinationState)]
    f("set_visits").argtypes = [vp, C.POINTER(JbVisitsDesc)]
    f("set_visit_probabilities").argtypes = [vp, C.c_double, C.c_double]
    f("flush_l2").argtypes = [vp, C.c_size_t]
    f("set_medical").argtypes = [vp, vp]
    f("set_halo_activities").argtypes = [vp, C.c_uint32]
    f("set_person_work_regions").argtypes = [vp, vp]
    f("set_variants").argtypes = [vp, C.c_uint32, vp, vp]
    f("mutate").argtypes = [vp, C.c_uint32, vp, C.c_uint64, C.c_uint32]
    f("fetch_health_events").argtypes = [vp, vp, vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    f("p2p_window").argtypes = [vp, vp]
    f("p2p_connect").argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    f("set_commute").argtypes = [vp, C.POINTER(JbCommuteDesc)]
    f("inject_commute").argtypes = [vp, vp]
    f("fetch_infection_events").argtypes = [vp, vp, vp, vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    for extra, args in (("set_halo_gids", [vp, vp, C.c_size_t]), ("halo_record_size_w", [vp]), ("stream_handle", [vp]),
                        ("timing_read_supergroups", [vp, dp, C.POINTER(C.c_uint64), C.c_int]),
                        ("set_concurrency", [vp, C.c_int])):
        if hasattr(lib, prefix + extra):
            getattr(lib, prefix + extra).argtypes = args
    if hasattr(lib, prefix + "halo_record_size_w"):
        getattr(lib, prefix + "halo_record_size_w").restype = C.c_size_t
    if hasattr(lib, prefix + "stream_handle"):
        getattr(lib, prefix + "stream_handle").restype = vp
    return lib


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load ``libjune_b200.so``.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(june_b200 has no CPU fallback)")
        _lib = bind(C.CDLL(LIB_PATH), "jb_")
        if _lib.jb_abi_version() != ABI_VERSION:
            raise RuntimeError("libjune_b200.so ABI version mismatch: rebuild")
    return _lib


def ptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
