#!/usr/bin/env python
"""End-to-end BPtrain throughput (Pfile on disk -> trained weights), host reader vs device-side reader, C2 shape.
   python scripts/gpu_reader_bench.py [n_sentences]      (run on the GPU box; writes its corpus under /tmp)"""
import importlib
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T = importlib.import_module("dnn-for-speech-enhancement_b200.tools.pfile")
EXE = os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "bin", "BPtrain")
LS = [2827, 2048, 2048, 2048, 257]


def main():
    n_sent = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    with tempfile.TemporaryDirectory(dir="/tmp") as d:
        t0 = time.time()
        feas, targs, mu, ivar = T.synth_corpus(n_sent, 257, 257, seed=1, min_len=200, max_len=400)
        T.write_pfile(f"{d}/fea.pfile", feas)
        T.write_pfile(f"{d}/targ.pfile", targs)
        T.write_norm(f"{d}/fea.norm", mu, ivar)
        frames = sum(f.shape[0] for f in feas)
        print(f"corpus: {n_sent} sentences, {frames} frames, generated in {time.time() - t0:.1f} s", flush=True)
        cv0 = n_sent - 20
        for tag, pf in (("host", 1), ("gpu", 1), ("host", 0), ("gpu", 0), ("host", 1), ("gpu", 1)):
            args = [f"fea_file={d}/fea.pfile", f"norm_file={d}/fea.norm", f"targ_file={d}/targ.pfile",
                    f"outwts_file={d}/{tag}.wts", f"log_file={d}/{tag}.log", "initwts_file=",
                    f"train_sent_range=0-{cv0 - 1}", f"cv_sent_range={cv0}-{n_sent - 1}", "fea_dim=257",
                    "fea_context=11", "targ_offset=5", "traincache=102400", "bunchsize=1024",
                    "layersizes=" + ",".join(map(str, LS)), "gpu_used=1", "init_randem_seed=7", "momentum=0.9",
                    "weightcost=0", "lrate=0.1", "dropoutflag=0", "visible_omit=0", "hid_omit=0", "nat=0",
                    f"reader={tag}", f"prefetch={pf}"]
            t0 = time.time()
            o = subprocess.run([EXE] + args, cwd=d, capture_output=True, text=True, timeout=900)
            wall = time.time() - t0
            log = open(f"{d}/{tag}.log").read()
            thr = re.search(r"Training throughput: (\d+) frames/sec", log)
            cv = re.search(r"CV over\. squared error: (\S+)", log)
            print(f"reader={tag} prefetch={pf}: exit {o.returncode}, wall {wall:.2f} s, training pass "
                  f"{thr.group(1) if thr else '?'} frames/s, CV {cv.group(1) if cv else '?'}", flush=True)
        same = open(f"{d}/host.wts", "rb").read() == open(f"{d}/gpu.wts", "rb").read()
        print("weights identical between readers:", same)


if __name__ == "__main__":
    main()
