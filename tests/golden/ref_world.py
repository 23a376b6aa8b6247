"""Build a small synthetic world out of the *reference's own* constructors
(``june.groups`` / ``june.geography`` / ``june.demography``), with no census data, and wire
a reference ``Simulator`` around it.

Golden-fixture tooling: runs only in the build container (needs ``/root/reference``), through
``oracle/ref_harness.py``.  Constructor cheat-sheet: SURVEY.md Appendix C.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_harness  # noqa: E402

REGION_NAMES = ["North East", "London"]
SECTORS = list("ABCDEFGHIJKLMNOPQRSTU")


def synthetic_rates_df():
    """Age-banded synthetic outcome-rate table in the reference's schema
    (``health_index.py:22-33``, ``disease_config.py:191``): index = closed 10-year
    ``pd.Interval``s, columns ``{gp,ch}_{param}_{male,female}``."""
    import pandas as pd

    bins = [(0, 9), (10, 19), (20, 29), (30, 39), (40, 49), (50, 59), (60, 69), (70, 79), (80, 89), (90, 99)]
    asym = [0.45, 0.40, 0.35, 0.30, 0.28, 0.25, 0.22, 0.20, 0.18, 0.15]
    mild = [0.50, 0.52, 0.52, 0.50, 0.45, 0.40, 0.35, 0.28, 0.22, 0.20]
    hosp = [0.004, 0.006, 0.02, 0.04, 0.06, 0.10, 0.16, 0.22, 0.26, 0.28]
    icu = [0.0005, 0.001, 0.004, 0.008, 0.015, 0.03, 0.05, 0.06, 0.05, 0.04]
    home_ifr = [0.0001, 0.0001, 0.0003, 0.0006, 0.001, 0.003, 0.008, 0.02, 0.04, 0.06]
    hosp_ifr = [0.0002, 0.0003, 0.001, 0.002, 0.005, 0.012, 0.03, 0.07, 0.12, 0.15]
    icu_ifr = [0.0001, 0.0002, 0.0008, 0.002, 0.004, 0.01, 0.02, 0.03, 0.03, 0.03]
    cols = {}
    for pop, mult in (("gp", 1.0), ("ch", 1.5)):
        for sex, smult in (("male", 1.1), ("female", 0.9)):
            for name, vals in (
                ("asymptomatic", asym), ("mild", mild), ("hospital", hosp), ("icu", icu),
                ("home_ifr", home_ifr), ("hospital_ifr", hosp_ifr), ("icu_ifr", icu_ifr),
            ):
                v = np.array(vals, dtype=float)
                if name not in ("asymptomatic", "mild"):
                    v = v * mult * smult
                cols[f"{pop}_{name}_{sex}"] = v
    idx = [pd.Interval(left=a, right=b, closed="both") for a, b in bins]
    return pd.DataFrame(cols, index=idx)


def build_world(n_people=3000, seed=0, with_hospital=True, with_care_homes=True,
                people_per_company=40, people_per_school=1000, n_super_areas=4, with_leisure=False, with_visits=False,
                with_commute=False):
    """Returns ``(world, disease_config)``.  People are created household by household so
    that person index order == household order."""
    ref_harness.setup()
    from june.demography import Person, Population
    from june.epidemiology.infection.disease_config import DiseaseConfig
    from june.geography import Area, Areas, Region, Regions, SuperArea, SuperAreas
    from june.global_context import GlobalContext
    from june.groups import (CareHome, CareHomes, Cemeteries, Companies, Company, Hospital,
                             Hospitals, Household, Households, School, Schools)
    from june.world import World

    rng = np.random.default_rng(seed)
    dc = DiseaseConfig("covid19")
    GlobalContext.set_disease_config(dc)

    regions = [Region(name=n) for n in REGION_NAMES]
    super_areas, areas = [], []
    for s in range(n_super_areas):
        reg = regions[s % len(regions)]
        sa = SuperArea(name=f"SA{s}", coordinates=(50.0 + s, 0.0), region=reg)
        reg.super_areas.append(sa)
        super_areas.append(sa)
        for a in range(4):
            ar = Area(name=f"A{s}_{a}", super_area=sa, coordinates=(50.0 + s, 0.1 * a),
                      socioeconomic_index=0.5)
            sa.areas.append(ar)
            areas.append(ar)

    # age pyramid (coarse, England-like)
    age_w = np.concatenate([np.full(18, 1.2), np.full(47, 1.3), np.full(20, 0.9), np.full(15, 0.25)])
    age_p = age_w / age_w.sum()
    hh_sizes = np.array([1, 2, 3, 4, 5, 6])
    hh_p = np.array([0.30, 0.34, 0.16, 0.13, 0.05, 0.02])

    population = Population()
    households, people = [], []
    while len(people) < n_people:
        area = areas[len(households) % len(areas)]
        size = int(rng.choice(hh_sizes, p=hh_p))
        hh = Household(type="family", area=area, max_size=8)
        # at least one adult
        ages = list(rng.choice(100, size=size, p=age_p))
        if max(ages) < 18:
            ages[0] = int(rng.integers(25, 60))
        for age in ages:
            p = Person.from_attributes(sex="m" if rng.random() < 0.5 else "f", age=int(age))
            area.add(p)
            hh.add(p, activity="residence")
            population.add(p)
            people.append(p)
        area.households.append(hh)
        households.append(hh)

    # care homes: move some 65+ into care homes (residence replaced)
    care_homes = []
    if with_care_homes:
        n_ch = max(1, n_people // 3000)
        old = [p for p in people if p.age >= 75]
        rng.shuffle(old)
        per = 20
        for c in range(n_ch):
            area = areas[(7 * c) % len(areas)]
            ch = CareHome(area=area, n_residents=per, n_workers=3)
            area.care_home = ch
            for p in old[c * per:(c + 1) * per]:
                # take the person out of their household
                hh = p.residence.group
                p.residence.people.remove(p)
                hh.residents = tuple(m for m in hh.residents if m is not p)
                ch.add(p, subgroup_type=ch.SubgroupType.residents, activity="residence")
                p.busy = False
            care_homes.append(ch)

    # schools: ages 4..18, teachers picked among adults
    schools = []
    n_schools = max(1, n_people // people_per_school)
    kids = [p for p in people if 4 <= p.age <= 18]
    adults = [p for p in people if 22 <= p.age <= 64]
    rng.shuffle(adults)
    adult_iter = iter(adults)
    specs = [(4, 11, "primary"), (11, 18, "secondary"), (4, 18, None)]
    for s in range(n_schools):
        amin, amax, sector = specs[s % len(specs)]
        area = areas[(3 * s) % len(areas)]
        sc = School(coordinates=area.coordinates, n_pupils_max=2000, age_min=amin, age_max=amax,
                    sector=sector, area=area)
        schools.append(sc)
    for i, k in enumerate(kids):
        cands = [sc for sc in schools if sc.age_min <= k.age <= sc.age_max]
        if not cands:
            continue
        sc = cands[i % len(cands)]
        sc.add(k)
    for sc in schools:
        n_pupils = sum(len(sg.people) for sg in sc.subgroups[1:])
        for _ in range(max(1, n_pupils // 20)):
            t = next(adult_iter)
            sc.add(t)  # older than age_max -> teachers (school.py:99-102)
    for sc in schools:
        sc.limit_classroom_sizes(25)
    for p in people:
        p.busy = False

    # hospital + care home workers
    hospitals = []
    if with_hospital:
        area = areas[0]
        h = Hospital(n_beds=200, n_icu_beds=20, area=area, coordinates=area.coordinates)
        for _ in range(max(3, n_people // 500)):
            w = next(adult_iter)
            h.add(w, subgroup_type=h.SubgroupType.workers)
            w.busy = False
        hospitals.append(h)
        for sa in super_areas:
            sa.closest_hospitals = [h]
    for ch in care_homes:
        for _ in range(3):
            w = next(adult_iter)
            ch.add(w, subgroup_type=ch.SubgroupType.workers, activity="primary_activity")
            w.busy = False

    # companies: remaining working-age adults
    companies = []
    n_comp = max(1, n_people // people_per_company)
    for c in range(n_comp):
        sa = super_areas[c % len(super_areas)]
        companies.append(Company(super_area=sa, n_workers_max=2000, sector=SECTORS[c % len(SECTORS)]))
    sizes = rng.lognormal(np.log(8), 1.0, n_comp)
    cp = sizes / sizes.sum()
    for w in adult_iter:
        if rng.random() < 0.85:
            c = companies[int(rng.choice(n_comp, p=cp))]
            c.add(w)
            w.busy = False
            w.sector = c.sector
    # lockdown status of everybody with a primary activity that is not a school pupil
    # (CloseCompanies / CloseSchools read it, individual_policies.py:497-510,651-770)
    lrng = np.random.default_rng(seed + 77)
    for p in people:
        if p.primary_activity is not None and p.age > 18:
            u = lrng.random()
            p.lockdown_status = "key_worker" if u < 0.25 else "furlough" if u < 0.35 else "random"

    world = World()
    world.areas = Areas(areas, ball_tree=False)
    world.super_areas = SuperAreas(super_areas, ball_tree=False)
    world.regions = Regions(regions)
    world.people = population
    world.households = Households(households)
    world.companies = Companies(companies)
    world.schools = Schools(schools)
    if hospitals:
        world.hospitals = Hospitals(hospitals, ball_tree=False)
    if care_homes:
        world.care_homes = CareHomes(care_homes)
    world.cemeteries = Cemeteries()
    if with_commute:
        # two cities of two super areas each; per city 2 city stations x 2 transports and one inter-city
        # station x 2 transports per super area (travel.py:125-131, city.py:83-102, station.py:52-76)
        from june.geography.city import Cities, City
        from june.geography.station import CityStation, InterCityStation
        from june.groups.travel import CityTransport, CityTransports, InterCityTransport, InterCityTransports
        from june.groups.travel.mode_of_transport import ModeOfTransport
        public, private = ModeOfTransport("Bus, minibus or coach", is_public=True), ModeOfTransport("Driving a car or van", is_public=False)
        cities, ctrans, itrans = [], [], []
        for c in range(max(1, n_super_areas // 2)):
            sas = super_areas[2 * c: 2 * c + 2]
            city = City(super_areas=[sa.name for sa in sas], super_area=sas[0], name=f"City{c}", coordinates=sas[0].coordinates)
            for sa in sas:
                sa.city = city
            for k in range(2):
                st = CityStation(city=city.name, super_area=sas[k % len(sas)])
                st.city_transports = [CityTransport(station=st) for _ in range(2)]
                ctrans += st.city_transports
                city.city_stations.append(st)
            for sa in sas:
                ist = InterCityStation(city=city.name, super_area=sa)
                ist.inter_city_transports = [InterCityTransport(station=ist) for _ in range(2)]
                itrans += ist.inter_city_transports
                city.inter_city_stations.append(ist)
            cities.append(city)
        for sa in super_areas:
            for city in cities:  # the station a commuter from this super area uses to reach `city`
                same = [st for st in city.inter_city_stations if st.super_area is sa]
                sa.closest_inter_city_station_for_city[city.name] = same[0] if same else city.inter_city_stations[0]
        crng = np.random.default_rng(seed + 99)
        for p in people:
            if p.primary_activity is not None and p.primary_activity.group.spec == "company":
                p.work_super_area = p.primary_activity.group.super_area
                p.mode_of_transport = public if crng.random() < 0.5 else private
                if p.mode_of_transport.is_public:
                    city = p.work_super_area.city
                    if p.area.super_area.city is city and crng.random() < 0.6:
                        city.internal_commuter_ids.add(p.id)
                    elif crng.random() < 0.8:
                        p.area.super_area.closest_inter_city_station_for_city[city.name].commuter_ids.add(p.id)
            else:
                p.mode_of_transport = private
        world.cities = Cities(cities, ball_tree=False)
        world.city_transports = CityTransports(ctrans)
        world.inter_city_transports = InterCityTransports(itrans)
    if with_leisure:
        # social venues: one pub per area, one grocery per two areas, one cinema per super area;
        # an area's candidates are hand-wired (the reference's distributor needs coordinates + a tree)
        from june.groups.leisure import Cinema, Cinemas, Groceries, Grocery, Pub, Pubs
        pubs = [Pub(area=a) for a in areas]
        groceries = [Grocery(area=a) for a in areas[::2]]
        cinemas = [Cinema(area=sa.areas[0]) for sa in super_areas]
        for i, a in enumerate(areas):
            a.social_venues = {
                "pub": tuple(pubs[j % len(pubs)] for j in (i, i + 1, i + 2)),
                "grocery": tuple(groceries[j % len(groceries)] for j in (i // 2, i // 2 + 1)),
                "cinema": (cinemas[(i // 4) % len(cinemas)],),
            }
        world.pubs = Pubs(pubs, make_tree=False)
        world.groceries = Groceries(groceries, make_tree=False)
        world.cinemas = Cinemas(cinemas, make_tree=False)
    if with_visits:
        # Household.residences_to_visit, as ResidenceVisitsDistributor.link_households_to_households / _to_care_homes
        # set it (residence_visits.py:189-296; hand-wired here: the reference's linker wants census household types):
        # two to three households of the same super area, and one household per care-home resident
        vrng = np.random.default_rng([seed, 4242])
        by_sa = {}
        for hh in households:
            by_sa.setdefault(id(hh.area.super_area), []).append(hh)
        for hh in households:
            pool = [h for h in by_sa[id(hh.area.super_area)] if h is not hh and h.n_residents > 0]
            if hh.n_residents == 0 or not pool:
                continue
            k = int(vrng.integers(2, 4))
            hh.residences_to_visit["household"] = tuple(pool[int(j)] for j in vrng.integers(0, len(pool), k))
        if with_care_homes:
            for ch in world.care_homes.members:
                pool = [h for h in by_sa.get(id(ch.area.super_area), []) if h.n_residents >= 2]
                for _ in ch.residents:
                    if pool:
                        h = pool[int(vrng.integers(0, len(pool)))]
                        h.residences_to_visit["care_home"] = (*h.residences_to_visit["care_home"], ch)
    return world, dc


VACCINES = {
    "Pfizer": dict(days_administered_to_effective=[2, 2, 5], days_effective_to_waning=[1, 2, 1], days_waning=[2, 2, 1],
                   waning_factor=0.8,
                   sterilisation_efficacies=[{"Covid19": {"0-100": 0.52}}, {"Covid19": {"0-70": 0.95, "70-100": 0.85}}, {"Covid19": {"0-100": 0.98}}],
                   symptomatic_efficacies=[{"Covid19": {"0-100": 0.6}}, {"Covid19": {"0-100": 0.9}}, {"Covid19": {"0-100": 0.98}}]),
    "AstraZeneca": dict(days_administered_to_effective=[3, 7, 5], days_effective_to_waning=[2, 1, 1], days_waning=[3, 1, 1],
                        waning_factor=1.0,
                        sterilisation_efficacies=[{"Covid19": {"0-100": 0.32}}, {"Covid19": {"0-100": 0.75}}, {"Covid19": {"0-100": 0.88}}],
                        symptomatic_efficacies=[{"Covid19": {"0-100": 0.4}}, {"Covid19": {"0-100": 0.75}}, {"Covid19": {"0-100": 0.88}}]),
}


def make_simulator(world, dc, total_days=10, cases_per_capita=0.01, with_policies=True,
                   initial_day="2020-03-02 0:00", schedule="default", with_leisure=False, with_visits=False, lockdown=False,
                   with_commute=False, vaccination=False):
    """Reference Simulator wired as ``example_scripts/run_simulation.py:332-512`` does, minus
    leisure / travel / records."""
    ref_harness.setup()
    from june.activity import ActivityManager
    from june.epidemiology.epidemiology import Epidemiology
    from june.epidemiology.infection import InfectionSelector, InfectionSelectors
    from june.epidemiology.infection.health_index.health_index import HealthIndexGenerator
    from june.epidemiology.infection_seed import InfectionSeed, InfectionSeeds
    from june.interaction import Interaction
    from june.policy import Hospitalisation, Policies, SevereSymptomsStayHome
    from june.simulator import Simulator
    from june.time import Timer

    hig = HealthIndexGenerator(dc, synthetic_rates_df())
    selector = InfectionSelector(dc, health_index_generator=hig)
    seed = InfectionSeed.from_uniform_cases(world=world, infection_selector=selector,
                                            cases_per_capita=cases_per_capita,
                                            date=initial_day.split(" ")[0],
                                            seed_past_infections=False)
    campaigns = None
    if vaccination:
        # two campaigns over disjoint target groups (so that np.random.choice is never needed):
        # the over-60s get doses 0 and 1 three days apart, care-home residents a single dose
        from june.epidemiology.vaccines.vaccination_campaign import VaccinationCampaign, VaccinationCampaigns
        from june.epidemiology.vaccines.vaccines import Vaccine
        pfizer = Vaccine.from_config_dict("Pfizer", VACCINES["Pfizer"])
        az = Vaccine.from_config_dict("AstraZeneca", VACCINES["AstraZeneca"])
        campaigns = VaccinationCampaigns([
            VaccinationCampaign(vaccine=pfizer, days_to_next_dose=[0, 3], dose_numbers=[0, 1], start_time="2020-03-03",
                                end_time="2020-03-08", group_by="age", group_type="60-100", group_coverage=0.6),
            VaccinationCampaign(vaccine=az, days_to_next_dose=[0], dose_numbers=[0], start_time="2020-03-04",
                                end_time="2020-03-07", group_by="age", group_type="30-45", group_coverage=0.5),
        ])
    epidemiology = Epidemiology(infection_selectors=InfectionSelectors([selector]),
                                infection_seeds=InfectionSeeds([seed]), vaccination_campaigns=campaigns)
    interaction = Interaction.from_file()
    if schedule == "default":
        # config_simulation.yaml:15-41 without commute/leisure activities
        wd_dur, we_dur = [1, 8, 1, 3, 11], [12, 12]
        wd_act = [["medical_facility", "residence"],
                  ["medical_facility", "primary_activity", "residence"],
                  ["medical_facility", "residence"],
                  ["medical_facility", "residence"],
                  ["medical_facility", "residence"]]
        we_act = [["medical_facility", "residence"], ["medical_facility", "residence"]]
    else:
        wd_dur, we_dur = [8, 16], [24]
        wd_act = [["medical_facility", "primary_activity", "residence"], ["medical_facility", "residence"]]
        we_act = [["medical_facility", "residence"]]
    timer = Timer(initial_day=initial_day, total_days=total_days,
                  weekday_step_duration=wd_dur, weekend_step_duration=we_dur,
                  weekday_activities=wd_act, weekend_activities=we_act)
    if with_commute:
        # config_simulation.yaml:15-41 with commute (no leisure)
        wd_dur, we_dur = [1, 8, 1, 3, 11], [12, 12]
        wd_act = [["medical_facility", "residence", "commute"],
                  ["medical_facility", "primary_activity", "residence"],
                  ["medical_facility", "residence", "commute"],
                  ["medical_facility", "residence"],
                  ["medical_facility", "residence"]]
        we_act = [["medical_facility", "residence"], ["medical_facility", "residence"]]
        timer = Timer(initial_day=initial_day, total_days=total_days,
                      weekday_step_duration=wd_dur, weekend_step_duration=we_dur,
                      weekday_activities=wd_act, weekend_activities=we_act)
    if with_leisure:
        # config_simulation.yaml:15-41 with leisure (no commute)
        wd_dur, we_dur = [1, 8, 1, 3, 11], [4, 4, 4, 12]
        wd_act = [["medical_facility", "residence"],
                  ["medical_facility", "primary_activity", "leisure", "residence"],
                  ["medical_facility", "residence"],
                  ["medical_facility", "leisure", "residence"],
                  ["medical_facility", "residence"]]
        we_act = [["medical_facility", "leisure", "residence"]] * 3 + [["medical_facility", "residence"]]
        timer = Timer(initial_day=initial_day, total_days=total_days,
                      weekday_step_duration=wd_dur, weekend_step_duration=we_dur,
                      weekday_activities=wd_act, weekend_activities=we_act)
    pol = [Hospitalisation(disease_config=dc), SevereSymptomsStayHome(start_time="1900-01-01", end_time="2100-01-01")] if with_policies else []
    if lockdown:
        from june.policy import CloseCompanies, CloseSchools, LimitLongCommute, Quarantine, Shielding
        for p in world.people:  # LimitLongCommute needs the work super area (set by the commute builder otherwise)
            if p.primary_activity is not None and p.primary_activity.group.spec == "company" and p.work_super_area is None:
                p.work_super_area = p.primary_activity.group.super_area
        llc = LimitLongCommute(start_time="2020-03-07", end_time="2100-01-01", apply_from_distance=150, going_to_work_probability=0.4)
        llc.initialize(world, None, None)
        pol += [
            Quarantine(start_time="2020-03-03", end_time="2100-01-01", n_days=7, n_days_household=14, compliance=0.8,
                       household_compliance=0.9),
            Shielding(start_time="2020-03-04", end_time="2100-01-01", min_age=70, compliance=0.6),
            CloseSchools(start_time="2020-03-05", end_time="2100-01-01", years_to_close=[5, 6, 7, 12, 13, 17],
                         attending_compliance=0.7),
            CloseCompanies(start_time="2020-03-06", end_time="2100-01-01", avoid_work_probability=0.3,
                           furlough_probability=0.2, key_probability=0.2),
            llc,
        ]
        CloseCompanies.initialize(world, None, None)
    policies = Policies(pol)
    a2sg = {
        "medical_facility": ["hospitals"] if world.hospitals is not None else [],
        "primary_activity": ["schools", "companies"],
        "residence": ["households"] + (["care_homes"] if world.care_homes is not None else []),
    }
    all_activities = ["medical_facility", "primary_activity", "residence"]
    leisure = None
    if with_leisure:
        from june.groups.leisure import generate_leisure_for_world
        a2sg["leisure"] = ["pubs", "cinemas", "groceries"] + (["household_visits", "care_home_visits"] if with_visits else [])
        all_activities = ["medical_facility", "primary_activity", "leisure", "residence"]
        leisure = generate_leisure_for_world(
            a2sg["leisure"], world,
            daytypes={"weekday": ["Monday", "Tuesday", "Wednesday", "Thursday", "Friday"], "weekend": ["Saturday", "Sunday"]})
    travel = None
    if with_commute:
        from june.groups.travel import Travel
        travel = Travel.__new__(Travel)  # only get_commute_subgroup is used (activity_manager.py:389)
        a2sg["commute"] = ["city_transports", "inter_city_transports"]
        all_activities = ["medical_facility", "commute", "primary_activity", "residence"]
    am = ActivityManager(world=world, policies=policies, timer=timer,
                         all_activities=all_activities, activity_to_super_groups=a2sg, leisure=leisure, travel=travel)
    sim = Simulator(world=world, interaction=interaction, timer=timer, activity_manager=am,
                    epidemiology=epidemiology, tracker=None, record=None)
    return sim
