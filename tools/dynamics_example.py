"""Wall time of the dynamics loop on the reference's example target (tests/golden copies of example/seq NMR + X-ray npz):
init_num initial decoys per model, then up to n_max decay -> fold iterations, both models' chains concurrently vs one after
the other, distogram update on the device vs on the host."""
import argparse, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trx2dyn
from trx2dyn import capi, dynamics, pipeline

ap = argparse.ArgumentParser()
ap.add_argument("--init-num", type=int, default=10)
ap.add_argument("--n-max", type=int, default=40)
a = ap.parse_args()
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
npzs = [os.path.join(G, "example_NMR.npz"), os.path.join(G, "example_Xray.npz")]
fasta = os.path.join(G, "example_seq.fasta")
for label, streams in (("warm-up", 2), ("both chains concurrently (2 streams)", 2), ("one chain after the other (1 stream)", 1)):
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.perf_counter()
        files = pipeline.run_single_from_npz("seq", fasta, npzs, tmp, init_num=a.init_num if label != "warm-up" else 2,
                                             n_max=a.n_max if label != "warm-up" else 2, seed=1, streams=streams)
        dt = time.perf_counter() - t0
    n1 = sum("conf_1_" in f for f in files)
    n2 = len(files) - n1
    print("%-42s %6.1f s  decoys: model 1 %d, model 2 %d  (%.2f s per decoy-iteration of the longer chain)" %
          (label, dt, n1, n2, dt / max(1, max(n1, n2) - a.init_num + 1)), flush=True)
