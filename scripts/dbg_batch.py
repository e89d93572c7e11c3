import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import fos_b200 as fos
from fos_b200 import problems
from test_gpu_batch import _load_batch, _nnls_batch
for B in (7, 12, 13, 40):
    plist = _nnls_batch(problems, B, 24, 30, scale=0.3, seed0=100)
    for ctas in (3, 1):
        H = _load_batch(fos, plist, batch_ctas=ctas, batch_hybrid=1)
        H.set_algorithm(fos.FISTA())
        done, st, recs, guess = H.solve_batch(300, 25, 1e-3)
        print("B", B, "ctas", ctas, "solve:", done.tolist()[:14], st.tolist()[:14], "rec0", ["%.3e" % v for v in recs[0][0, 1:8]], flush=True)
