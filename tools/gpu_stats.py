import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from centroflye_b200.engine import Engine
eng = Engine("cuda:0")
unit, batch, units = bench.make_inputs(float(sys.argv[1]) if len(sys.argv) > 1 else 1.0)
lo, hi = bench.band()
P = bench.PARAMS
reads, dunits = eng.upload_reads(batch, P["k"]), eng.upload_units(units, P["k"])
index, csr, res = eng.recruit(reads, dunits, P["k"], lo, hi, P["max_nonuniq"], P["min_d"], P["max_d"], P["min_coverage"])
c = eng.last_pair_counters
print("cloud entries", csr.n_entries, "units", csr.n_units, "rare", index.n)
print("counters", c)
print(f"passes {c[4]}, windows {c[5]}, keys {c[7]}, lane-iters {c[6]}, increments {c[2]}, splits {c[3]}")
if c[5]:
    print(f"keys/window {c[7]/c[5]:.1f}  lane-iters/key {c[6]/max(c[7],1):.2f}  windows/pass {c[5]/max(c[4],1):.1f} keys/pass {c[7]/max(c[4],1):.0f}")
