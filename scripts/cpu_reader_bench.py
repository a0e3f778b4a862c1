#!/usr/bin/env python
"""Host-side reader throughput on the machine at hand (no GPU needed): synthetic C2-shaped corpus (257-dim features and
targets, 11-frame context), 102 400-sample chunks, reader=host vs reader=gpu planner (bin/reader_bench).
   python scripts/cpu_reader_bench.py [n_sentences]   > profiles/<round>_host_reader.log"""
import importlib
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T = importlib.import_module("dnn-for-speech-enhancement_b200.tools.pfile")
EXE = os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "bin", "reader_bench")


def main():
    n_sent = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
    with tempfile.TemporaryDirectory(dir="/tmp") as d:
        feas, targs, mu, ivar = T.synth_corpus(n_sent, 257, 257, seed=1, min_len=200, max_len=400)
        T.write_pfile(f"{d}/fea.pfile", feas)
        T.write_pfile(f"{d}/targ.pfile", targs)
        T.write_norm(f"{d}/fea.norm", mu, ivar)
        print(f"corpus: {n_sent} sentences, {sum(f.shape[0] for f in feas)} frames; host threads: {os.cpu_count()}")
        args = [f"fea_file={d}/fea.pfile", f"norm_file={d}/fea.norm", f"targ_file={d}/targ.pfile",
                f"outwts_file={d}/o.wts", f"log_file={d}/o.log", "initwts_file=", f"train_sent_range=0-{n_sent - 21}",
                f"cv_sent_range={n_sent - 20}-{n_sent - 1}", "fea_dim=257", "fea_context=11", "targ_offset=5",
                "traincache=102400", "bunchsize=1024", "layersizes=2827,2048,2048,2048,257", "gpu_used=1",
                "init_randem_seed=7", "momentum=0.9", "weightcost=0", "lrate=0.1", "dropoutflag=0", "visible_omit=0",
                "hid_omit=0", "nat=0"]
        for reader in ("host", "gpu"):
            out = subprocess.run([EXE] + args + [f"reader={reader}"], cwd=d, capture_output=True, text=True).stdout
            rows = [l for l in out.splitlines() if l.startswith("rep 2")]
            print(f"reader={reader} (third pass, page cache warm):")
            for l in rows:
                print("   ", l)


if __name__ == "__main__":
    main()
