#!/usr/bin/env python
"""TEST INFRASTRUCTURE - records the schema of the reference's own .vul output (Output.save_out, op.py:3216-3255): the
unmodified reference runs `--steps` steps of a staged config, writes its pickle, and the key / type / shape tree goes to
tests/golden/<cfg>_vul_schema.json (the pickle itself is ~100 MB and is not kept).
usage: PYTHONHASHSEED=0 python oracle/make_vul_schema.py --config HD189 --steps 3"""
import argparse, json, os, pickle, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def describe(v, depth=0):
    if isinstance(v, np.ndarray):
        return {"type": "ndarray", "dtype": str(v.dtype), "shape": list(v.shape)}
    if isinstance(v, dict):
        keys = list(v.keys())
        d = {"type": "dict", "n": len(keys), "key_type": type(keys[0]).__name__ if keys else None}
        if depth < 1 and keys and all(isinstance(k, str) for k in keys):
            d["items"] = {k: describe(v[k], depth + 1) for k in keys}
        elif keys:
            d["first_value"] = describe(v[keys[0]], depth + 1)
        return d
    if isinstance(v, (list, tuple)):
        return {"type": type(v).__name__, "n": len(v), "first": describe(v[0], depth + 1) if len(v) else None}
    return {"type": type(v).__name__}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="HD189")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    sys.path.insert(0, HERE)
    import ref_session
    refdir = "/tmp/vulcan_ref_%s" % a.config
    s = ref_session.setup(refdir)
    s.cfg.count_max = a.steps
    s.cfg.out_name = "schema_probe.vul"
    s.integ(s.var, s.atm, s.para, s.make_atm)
    s.output.save_out(s.var, s.atm, s.para, refdir)
    path = os.path.join(refdir, s.cfg.output_dir, s.cfg.out_name)
    with open(path, "rb") as f:
        data = pickle.load(f)
    schema = {sec: {k: describe(v) for k, v in data[sec].items()} for sec in ("variable", "atm", "parameter")}
    schema["_top"] = list(data.keys())
    out = os.path.join(GOLD, "%s_vul_schema.json" % a.config)
    with open(out, "w") as f:
        json.dump(schema, f, indent=1, sort_keys=True)
    os.remove(path)
    print("wrote", out, {k: len(v) for k, v in schema.items()})
