"""Pretty-print a bench.py JSON line (headline + modes block + comm_check)."""
import json
import sys

for path in sys.argv[1:]:
    for l in open(path):
        if not l.startswith("{"):
            continue
        d = json.loads(l)
        if "unavailable" in d:
            print(d)
            continue
        e2e = (d.get("e2e") or {}).get("ms_per_step")
        print(f"{d.get('impl')} {d['config']['model']} {d['config']['parallelism']}: {d['ms_per_step']:.3f} ms/step, {d['value']:.0f} tok/s, "
              f"e2e {e2e and round(e2e, 3)} ms, exposed {d.get('exposed_comm_ms_per_step')}, peak {d['peak_hbm_bytes'] / 2**30:.2f} GiB, "
              f"loss {d['final_loss']:.3f}, clocks {d['clocks'].get('sm_mhz')} MHz x{d['clocks'].get('samples')} {d['clocks'].get('reasons')}")
        cc = d.get("comm_check")
        if cc:
            print("  comm_check:", {k: v for k, v in cc.items() if k != "cases"})
        for k, v in (d.get("modes") or {}).items():
            if "error" in v:
                print(f"  {k}: ERROR {v['error']}")
            else:
                print(f"  {k}: {v['ms_per_step']:.3f} ms/step, {v['value']:.0f} tok/s, peak {v['peak_hbm_bytes'] / 2**30:.2f} GiB, "
                      f"exposed {v.get('exposed_comm_ms_per_step')}, loss {v['final_loss']:.3f}")
