#!/usr/bin/env python
"""Print the few numbers of a bench.py JSON line one looks at when comparing variants (stdin or file)."""
import json, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 and sys.argv[1] != "-" else sys.stdin.read()
tag = sys.argv[2] if len(sys.argv) > 2 else ""
d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
k = d.get("kernel_ms", {})
print(tag, "value %.4g" % d["value"], "step", d.get("step_ms"), "K1 %.3f solve %.3f" % (k.get("geometry(K1 incl. pack+dPdrho)", float("nan")), k.get("solve(K2+K3+argmax)", float("nan"))),
      "e2e %.4g" % d.get("e2e", {}).get("value", float("nan")), "iters", d["config"].get("mean_solver_iterations"), "bad", d["config"].get("bad_solves"))
