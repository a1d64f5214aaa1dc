"""Print the few numbers of a bench.py JSON line (stdin) that matter when comparing library variants."""
import json, sys
for ln in sys.stdin:
    if ln.startswith("{"):
        d = json.loads(ln)
        print(sys.argv[1] if len(sys.argv) > 1 else "", "value %.2f ms/step %.2f e2e %s us/it %.3f iterations %d" % (
            d["value"], d["ms_per_step"], d["e2e"]["value"] if d.get("e2e") else None, d["roofline"]["us_per_launch"],
            d["roofline"]["launches"]))
