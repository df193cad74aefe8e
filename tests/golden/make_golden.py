"""Regenerates the committed fixtures from the read-only reference checkout.

Run in the build container only (`/root/reference` does not exist on the GPU box):
    python tests/golden/make_golden.py

Writes
  tests/golden/step_kat.json          the reference's own known-answer traces
                                      (pymc3/tests/test_step.py:163-266 HamiltonianMC,
                                      :371-474 NUTS; setup in check_trace :488-527)
  tests/golden/devguide_logp.json     docs/source/developer_guide.rst:715-737 (+ :151-155)
  pymc3_b200/data/sp500_log_returns.npy   log-returns of examples/data/SP500.csv as built
                                      in docs/source/notebooks/stochastic_volatility.ipynb cell 6
  pymc3_b200/data/radon_county_counts.json  rows per county in examples/data/radon.csv
The reference cannot be imported (no Theano), so test_step.py is parsed with `ast`.
"""
import ast
import json
import os
import re

import numpy as np
import pandas as pd

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def kat_vectors():
    src = open(os.path.join(REF, "pymc3/tests/test_step.py")).read()
    tree = ast.parse(src)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "master_samples":
            for key, val in zip(node.value.keys, node.value.values):
                arr = ast.literal_eval(val.args[0])
                out[key.id] = [float(v) for v in arr]
    assert len(out["NUTS"]) == 100 and len(out["HamiltonianMC"]) == 100
    return out


def devguide():
    txt = open(os.path.join(REF, "docs/source/developer_guide.rst")).read()
    blk = txt[txt.index("inputlist = [np.random.randn"):]
    blk = blk[blk.index(".. parsed-literal::"):blk.index("class") if "class" in blk[:10] else 4000]
    nums = lambda s: [float(x) for x in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", s)]
    z = nums(blk[blk.index("'z': array("):blk.index("'x': array(")])
    x = nums(blk[blk.index("'x': array("):blk.index("])}") + 2])
    tail = blk[blk.index("(array("):]
    tail = tail[:tail.index("]))") + 3]
    t = nums(tail)
    assert len(z) == 10 and len(x) == 10 and len(t) == 21, (len(z), len(x), len(t))
    return {"z": z, "x": x, "logp": t[0], "dlogp": t[1:],
            "scalar_model": {"z": 2.5, "x_logp": -4.0439386, "model_logp": -6.6973152}}


def main():
    with open(os.path.join(HERE, "step_kat.json"), "w") as f:
        json.dump(kat_vectors(), f, indent=0)
    with open(os.path.join(HERE, "devguide_logp.json"), "w") as f:
        json.dump(devguide(), f, indent=0)
    sp = pd.read_csv(os.path.join(REF, "pymc3/examples/data/SP500.csv"), index_col="Date")
    change = np.log(sp["Close"]).diff().dropna().to_numpy(dtype="d")
    np.save(os.path.join(ROOT, "pymc3_b200/data/sp500_log_returns.npy"), change)
    rd = pd.read_csv(os.path.join(REF, "pymc3/examples/data/radon.csv"))
    counts = np.bincount(rd.county_code.values).tolist()
    with open(os.path.join(ROOT, "pymc3_b200/data/radon_county_counts.json"), "w") as f:
        json.dump({"counts": counts, "floor_rate": float(rd.floor.mean()),
                   "n_rows": int(len(rd))}, f)
    print("T =", len(change), "counties =", len(counts), "rows =", len(rd))


if __name__ == "__main__":
    main()
