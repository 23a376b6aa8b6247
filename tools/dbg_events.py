import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
from test_events import *
w = generate_world(60_000, seed=3, sector_betas=load_defaults()["company_sector_betas"])
(a, b), beta = engines(w, ["cuda", "oracle"])
grp_sg = w.group_supergroup()
for i, ((ba, ea), (bb, eb)) in enumerate(zip(run(a, w, beta, 8), run(b, w, beta, 8))):
    for key in ("infected", "group", "variant", "infector"):
        if not np.array_equal(ea[key], eb[key]):
            bad = np.nonzero(ea[key] != eb[key])[0] if len(ea[key]) == len(eb[key]) else []
            print("step", i, key, "mismatch", len(ea[key]), len(eb[key]), "n_bad", len(bad))
            for q in bad[:8]:
                g = ea["group"][q]
                print("   infected", ea["infected"][q], "group", g, w.supergroups[grp_sg[g]].name, "size", w.grp_offset[g+1]-w.grp_offset[g],
                      "cuda infector", ea["infector"][q], "oracle infector", eb["infector"][q],
                      "trans", ba["trans_prob"][ea["infector"][q]] if ea["infector"][q] < w.n_people else None, bb["trans_prob"][eb["infector"][q]])
print("done")
