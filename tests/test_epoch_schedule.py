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
