"""Host-side mirrors against values dumped from the running reference (tests/golden/*.npz,
SURVEY.md Appendix B)."""
import datetime

import numpy as np
import pytest

from june_b200.disease import DiseaseConfig, outcome_probabilities, synthetic_rates
from june_b200.interaction import Interaction, processed_school_matrix
from june_b200.policy import Policies, SocialDistancing, MaskWearing
from june_b200.timer import Timer
from june_b200.world import BetaRule

from helpers import Golden


def test_raw_contact_matrices_match_reference():
    g = Golden("w3k")
    inter = Interaction.from_file()
    for k in g.z.files:
        if k.startswith("cm_"):
            spec = k[3:]
            np.testing.assert_array_equal(inter.contact_matrices[spec], g.z[k], err_msg=spec)
    ref = {str(k): float(v) for k, v in zip(g.z["beta_keys"], g.z["beta_vals"])}
    assert ref == {k: float(v) for k, v in inter.betas.items()}


def test_survey_appendix_b_spot_values():
    inter = Interaction.from_file()
    assert inter.contact_matrices["company"][0, 0] == 15.408000000000001
    np.testing.assert_allclose(inter.contact_matrices["household"][2], [4.4268, 2.5984, 3.86208, 3.86208], rtol=1e-15)
    s = inter.contact_matrices["school"]
    assert s.shape == (32, 32) and s[0, 0] == 15.75 and s[1, 1] == 8.625
    assert s[0, 1] == pytest.approx(48.60000000000001, rel=1e-15) and s[5, 9] == pytest.approx(0.06986249999999998, rel=1e-14)
    m = processed_school_matrix(s, (4, 5, 6))
    exp = [[15.75, 16.2, 16.2, 16.2], [0.81, 8.625, 2.5875, 0.77625], [0.81, 2.5875, 8.625, 2.5875], [0.81, 0.77625, 2.5875, 8.625]]
    np.testing.assert_allclose(m, exp, rtol=1e-14)


def test_outcome_table_matches_reference_health_index():
    g = Golden("w3k")
    dc = DiseaseConfig("covid19")
    rates, bins = synthetic_rates()
    prob = outcome_probabilities(dc, rates, bins)
    np.testing.assert_allclose(prob, g.outcome_prob, rtol=1e-14, atol=0)


def test_processed_beta_rules():
    inter = Interaction.from_file()
    inter.beta_reductions = {"company": 0.5, "school": 0.8, "household": 0.9}
    # interactive.py:154-177
    assert inter.processed_beta(BetaRule("company"), 0.95, None) == 0.371 * (1 + 0.95 * 1.0 * (0.5 - 1))
    assert inter.processed_beta(BetaRule("company"), 0.95, 4) == 0.371 * (1 + 0.95 * 0.5 * (0.5 - 1))
    # school.py:487-522: key falls back to "school" for beta and reduction
    assert inter.processed_beta(BetaRule("primary_school", "school"), 1.0, None) == 0.07 * (1 + 1.0 * 1.0 * (0.8 - 1))
    # household.py:305-307: no tier term
    assert inter.processed_beta(BetaRule("household", None, False), 1.1, 4) == 0.208 * (1 + 1.1 * (0.9 - 1))


def test_interaction_policies_multiply():
    p = Policies([SocialDistancing("2020-03-16", "2020-03-24", beta_factors={"pub": 0.5, "company": 0.8}),
                  MaskWearing("2020-03-20", "2020-04-01", compliance=0.5, beta_factor=0.5, mask_probabilities={"company": 0.4})])
    assert dict(p.beta_reductions(datetime.datetime(2020, 3, 15))) == {}
    assert dict(p.beta_reductions(datetime.datetime(2020, 3, 17))) == {"pub": 0.5, "company": 0.8}
    r = p.beta_reductions(datetime.datetime(2020, 3, 21))
    assert r["company"] == 0.8 * (1 - 0.4 * 0.5 * 0.5)


def test_default_policy_file_parses():
    p = Policies.from_file()
    names = {type(x).__name__ for x in p.policies}
    assert {"Hospitalisation", "SevereSymptomsStayHome", "RegionalCompliance", "SocialDistancing", "MaskWearing"} <= names
    assert p.step_flags(datetime.datetime(2020, 3, 1)) == 0x3


def test_timer_schedule_like_reference():
    # june/time.py:110-161 semantics: shift resets when the calendar day changes
    t = Timer(initial_day="2020-03-06 0:00", total_days=3, weekday_step_duration=[8, 16], weekend_step_duration=[24],
              weekday_activities=[["primary_activity", "residence"], ["residence"]], weekend_activities=[["residence"]])
    seen = []
    while t.date < t.final_date:
        seen.append((t.date.day, t.shift, t.duration, tuple(t.activities), t.day_type))
        next(t)
    assert seen == [(6, 0, 8 / 24, ("primary_activity", "residence"), "weekday"), (6, 1, 16 / 24, ("residence",), "weekday"),
                    (7, 0, 1.0, ("residence",), "weekend"), (8, 0, 1.0, ("residence",), "weekend")]
    assert t.now == 3.0


def test_summary_csv_rows(tmp_path):
    """Per-region daily rows (summary.csv schema) from the per-step device summaries."""
    from june_b200.interaction import load_defaults
    from june_b200.simulator import Epidemiology, InfectionSeed, Simulator
    from june_b200.synth import generate_world
    from helpers import make_engine, simulator_over
    w = generate_world(12_000, seed=2)
    inter = Interaction.from_file()
    dc = DiseaseConfig("covid19")
    rates, bins = synthetic_rates()
    prob = outcome_probabilities(dc, rates, bins)
    cfg = load_defaults()["simulation"]
    cfg["time"]["total_days"] = 8
    eng = make_engine("oracle", w, prob, interaction=inter, disease=dc)
    epi = Epidemiology(infection_seeds=[InfectionSeed(0.02, cfg["time"]["initial_day"])])
    sim = simulator_over(eng).from_file(w, inter, policies=Policies.from_file(disease_config=dc), epidemiology=epi,
                                        disease_config=dc, outcome_prob=prob, record=True)
    sim.timer = Timer.from_config(cfg)
    sim.activity_manager.timer = sim.timer
    sim.run()
    rows = sim.summary_rows()
    # one row per time step and region with anything to report (records_writer.py:764-778)
    assert 8 * len(w.region_names) < len(rows) <= len(sim.history) * len(w.region_names)
    assert set(rows[0]) >= set(Simulator.SUMMARY_HEADER)
    per_day = {}
    for r in rows:
        per_day.setdefault(r["time_stamp"], 0)
        per_day[r["time_stamp"]] += r["daily_infected"]
    tot = sum(h["n_new_infections"] for h in sim.history)
    n_seeds = int(round(0.02 * w.n_people))  # the seeds are rows of the infections table too (infection_seed.py:196-205)
    assert sum(per_day.values()) == tot + n_seeds and tot > 0
    sim.write_summary_csv(str(tmp_path / "summary.csv"))
    # the reference's header, column for column (records_writer.py:151-164)
    assert (tmp_path / "summary.csv").read_text().splitlines()[0] == (
        "time_stamp,region,current_infected,daily_infected,current_hospitalised,daily_hospitalised,"
        "current_intensive_care,daily_intensive_care,daily_hospital_deaths,daily_deaths")


def test_policy_file_with_a_policy_that_is_not_evaluated_is_refused_for_its_window(tmp_path):
    """A policy key this package does not evaluate must not silently vanish from a run in which it would be active."""
    import datetime
    import yaml
    from june_b200.policy import Policies
    path = tmp_path / "policy.yaml"
    path.write_text(yaml.safe_dump({
        "social_distancing": {"start_time": "2020-03-16", "end_time": "2020-03-24", "beta_factors": {"pub": 0.5}},
        "school_quarantine": {"start_time": "2020-04-01", "end_time": "2020-05-01", "compliance": 0.5, "n_days": 7},
    }))
    pol = Policies.from_file(str(path))
    assert [p.policy_type for p in pol.policies] == ["interaction"] and [k for k, _, _ in pol.not_evaluated] == ["school_quarantine"]
    pol.check_window(datetime.datetime(2020, 3, 1), datetime.datetime(2020, 4, 1))  # ends when the policy starts: fine
    with pytest.raises(NotImplementedError, match="school_quarantine"):
        pol.check_window(datetime.datetime(2020, 3, 1), datetime.datetime(2020, 4, 2))
    assert Policies.from_file(str(path), strict=False).not_evaluated == []
    Policies.from_file().check_window(datetime.datetime(2020, 3, 1), datetime.datetime(2021, 3, 1))  # the shipped defaults


def test_leisure_tables_vector_path_equals_per_person_calls():
    """``Leisure.generate_leisure_probabilities_for_timestep`` works on the 100 ages of a (region, sex) at once; a
    distributor subclass that overrides only the per-person ``get_poisson_parameter`` takes the per-person path.
    Both must give the same tables, bit for bit (the arithmetic is the reference's, ``leisure.py:205-228``,
    ``social_venue_distributor.py:176-205``)."""
    import datetime
    from june_b200.leisure import SEXES, Leisure, PubDistributor, GroceryDistributor

    class PerPersonPub(PubDistributor):
        def get_poisson_parameter(self, sex, age, day_type, working_hours, **kw):  # same numbers, scalar path only
            return super().get_poisson_parameter(sex=sex, age=age, day_type=day_type, working_hours=working_hours, **kw)

    regions = ["North", "South"]
    fast = Leisure({"pub": PubDistributor.from_config(), "grocery": GroceryDistributor.from_config()}, region_names=regions)
    slow = Leisure({"pub": PerPersonPub.from_config(), "grocery": GroceryDistributor.from_config()}, region_names=regions)
    red = {"pub": {dt: {s: [0.25 + 0.005 * a for a in range(100)] for s in ("m", "f")} for dt in ("weekday", "weekend")}}
    for L in (fast, slow):
        L.regional_compliance = {"North": 0.8, "South": 1.0}
        L.policy_reductions = red
        L.closed_venues = {"South": {"grocery"}}
    for date in (datetime.datetime(2020, 3, 31, 19), datetime.datetime(2020, 4, 4, 13)):
        for working_hours in (False, True):
            a = fast.generate_leisure_probabilities_for_timestep(3 / 24, working_hours, date)
            b = slow.generate_leisure_probabilities_for_timestep(3 / 24, working_hours, date)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            assert a[0].max() > 0 and (a[1][1, :, :, 1] == a[1][1, :, :, 0]).all(), "groceries closed in the South"
    view = fast.probabilities_by_region_sex_age
    row = view["North"]["f"][40]
    assert row["does_activity"] == a[0][0, list(SEXES).index("f"), 40] and set(row["activities"]) == {"pub", "grocery"}


def test_step_stats_as_dict_and_bench_kernel_names():
    from june_b200 import _lib
    st = _lib.JbStepStats()
    st.n_new_infections, st.n_halo_infected = 3, 11
    st.n_by_activity[0], st.n_by_activity[4] = 5, 9
    d = st.as_dict()
    assert d["n_new_infections"] == 3 and d["n_halo_infected"] == 11 and d["n_by_activity"] == [5, 0, 0, 0, 9]
    assert set(d) == {n for n, _ in _lib.JbStepStats._fields_}
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    text = ("supergroup 2: groups 9, activity 2, dynamic 0, mirror (slot-ordered state), tiles 5, span items 7, warp-list 0, cta-list 3, mean tiled group 1.00\n"
            "supergroup 3: groups 99, activity 4, dynamic 0, stream tiles (slot == person), tiles 40, span items 0, warp-list 0, cta-list 0, mean tiled group 2.34\n"
            "supergroup 4: groups 5, activity 4, dynamic 0, gather tiles, tiles 0 (none valid), span items 0, warp-list 5, cta-list 0, mean tiled group 0.00\n")
    assert bench.kernels_by_supergroup(text) == {2: "k_span + k_groups", 3: "k_tiles", 4: "k_groups"}
