import sys, os, json
sys.path.insert(0, os.getcwd())
from qgtc_ppopp22_b200 import main_qgtc as drv, cluster as C
g = C.make_synthetic_graph("ppi", 1500)
base = ["--dataset", "ppi", "--n-classes", "121", "--bit_width", "2", "--use_QGTC", "--n-epochs", "20"]
for k in (8, 16, 25, 38, 75):
    a = drv.build_parser().parse_args(base + ["--cuda-graph", "--streams", str(k)])
    r = drv.run(a, graph=g, quiet=True)
    print(json.dumps({"streams": k, "epoch_ms": round(r["epoch_ms"], 4)}), flush=True)
