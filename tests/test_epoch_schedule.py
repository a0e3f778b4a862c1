"""CPU test of the in-process epoch loop's schedule: the momentum each epoch gets must be, to the bit,
the float the reference binary would parse from the command line the reference's Perl driver builds
(finetune_DNN_speech_enhancement_dropout_NAT.pl:65,137,219: 0.5, then += 0.04 per epoch up to epoch 10, then 0.9) —
perl itself does the arithmetic and the number -> string conversion here."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "bin", "epoch_dump")

PERL = r'''
my $momentum = 0.5; print "momentum=$momentum\n";
for (my $i = 2; $i <= 10; $i++) { $momentum = $momentum + 0.04; print "momentum=$momentum\n"; }
for (my $i = 11; $i <= 14; $i++) { $momentum = 0.9; print "momentum=$momentum\n"; }
'''


@pytest.mark.skipif(not os.path.exists(TOOL), reason="epoch_dump not built")
@pytest.mark.skipif(shutil.which("perl") is None, reason="perl not installed")
def test_momentum_schedule_matches_perl_driver_bit_for_bit():
    perl = subprocess.run(["perl", "-e", PERL], capture_output=True, text=True, check=True).stdout.split()
    want = [np.float32(float(tok.split("=")[1])) for tok in perl]       # atof + narrowing, as Interface.cc:176
    out = subprocess.run([TOOL, "0.5", "0.04", "0.9", str(len(want)), "models/mlp.%d.wts"], capture_output=True,
                         text=True, check=True).stdout.strip().splitlines()
    assert len(out) == len(want)
    for e, (line, w) in enumerate(zip(out, want), start=1):
        num, bits, name = line.split()
        assert int(num) == e and name == f"models/mlp.{e}.wts"
        assert int(bits, 16) == int(np.array(w).view(np.uint32)), (e, bits, w)


@pytest.mark.skipif(not os.path.exists(TOOL), reason="epoch_dump not built")
def test_constant_momentum_and_plain_names():
    out = subprocess.run([TOOL, "0.9", "0", "0.9", "3", "plain.wts"], capture_output=True, text=True,
                         check=True).stdout.strip().splitlines()
    bits = {l.split()[1] for l in out}
    assert bits == {format(int(np.array(np.float32(0.9)).view(np.uint32)), "08x")}
    assert all(l.split()[2] == "plain.wts" for l in out)


REF_PL = "/root/reference/finetune_DNN_speech_enhancement_dropout_NAT.pl"
STUB = r'''#!/bin/bash
echo "$@" >> "$(dirname "$0")/../calls.txt"
for a in "$@"; do case "$a" in outwts_file=*) : > "${a#outwts_file=}";; esac; done
exit 1
'''


@pytest.mark.skipif(not os.path.exists(REF_PL), reason="reference Perl driver not present (GPU box)")
@pytest.mark.skipif(not os.path.exists(TOOL), reason="epoch_dump not built")
@pytest.mark.skipif(shutil.which("perl") is None, reason="perl not installed")
def test_unmodified_perl_driver_command_lines():
    """The reference's Perl epoch driver, UNMODIFIED and run by perl from where it lies, against a stub binary that
    records its command lines (success exit status 1, like BPtrain.cc:100).  All 100 invocations must (a) use only keys
    our Interface acts on (`numlayers` is ignored by the reference too, Interface.cc:227-243), and (b) follow, epoch for
    epoch, the schedule the in-process loop (epochs=N) applies: momentum to the bit, init_randem_seed + 345 per epoch,
    mlp.<i>.wts / .log names, initwts = the previous epoch's output."""
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        exe_dir = os.path.join(d, "code_BP_GPU_DNN_Dropout_NAT_speech_enhancement_GPU1")   # .pl:29
        os.mkdir(exe_dir)
        stub = os.path.join(exe_dir, "BPtrain")
        open(stub, "w").write(STUB)
        os.chmod(stub, 0o755)
        subprocess.run(["perl", REF_PL], cwd=d, capture_output=True, text=True, timeout=120, check=True)
        calls = [dict(tok.split("=", 1) for tok in line.split()) for line in open(os.path.join(d, "calls.txt"))]
    assert len(calls) == 100
    src = open(os.path.join(ROOT, "dnn-for-speech-enhancement_b200", "host", "Interface.cc")).read()
    for key in calls[0]:
        assert key == "numlayers" or f'"{key}"' in src, f"the Perl driver passes {key}=, which our parser does not know"
    first = calls[0]
    n = len(calls)
    sched = subprocess.run([TOOL, first["momentum"], "0.04", "0.9", str(n), "DIR/mlp.%d.wts"], capture_output=True,
                           text=True, check=True).stdout.strip().splitlines()
    mlp_dir = os.path.dirname(first["outwts_file"])
    for e, (c, line) in enumerate(zip(calls, sched)):
        num, bits, name = line.split()
        assert int(bits, 16) == int(np.array(np.float32(float(c["momentum"]))).view(np.uint32)), (e + 1, c["momentum"])
        assert c["outwts_file"] == name.replace("DIR", mlp_dir) and c["log_file"] == f"{mlp_dir}/mlp.{e + 1}.log"
        assert int(c["init_randem_seed"]) == int(first["init_randem_seed"]) + 345 * e          # seed_step default
        if e > 0:
            assert c["initwts_file"] == calls[e - 1]["outwts_file"]
        for key in ("lrate", "bunchsize", "layersizes", "dropoutflag", "visible_omit", "hid_omit", "traincache"):
            assert c[key] == first[key]
