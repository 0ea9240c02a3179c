"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/figbird_cuda.h declares,
its host-side helpers agree with the oracle, and the product path FAILS LOUDLY when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    hdr = open(os.path.join(ROOT, "include", "figbird_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fb_[a-z_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(built):
    from gapprober_b200 import figbird
    L = figbird.load_library()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"libfigbird_cuda.so does not export {n}"
    assert sorted(figbird.ABI_SYMBOLS) == names, "figbird.ABI_SYMBOLS out of sync with the header"
    assert L.fb_abi_version() == 1


def test_library_holds_sm100a_code(built):
    from gapprober_b200 import figbird
    out = subprocess.run(["cuobjdump", "-lelf", figbird.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out


def test_gap_range_matches_oracle(built):
    from gapprober_b200 import gap_range
    for g in list(range(1, 1200)) + [2500, 4999, 5000, 5001]:
        assert gap_range(g) == O.gap_range(g), g


def test_partition_plan_is_lpt(built):
    from gapprober_b200.figbird import partition_plan
    rng = np.random.default_rng(0)
    cost = rng.uniform(1, 100, size=200)
    for world in (1, 2, 4, 8):
        own = partition_plan(cost, world)
        assert own.min() >= 0 and own.max() < world
        loads = np.array([cost[own == r].sum() for r in range(world)])
        assert loads.max() - loads.min() <= cost.max() + 1e-9      # LPT guarantee
    assert list(partition_plan([5, 3, 3, 2, 2, 1], 2)) == [0, 1, 1, 0, 1, 0]


def test_no_cpu_fallback(built):
    """Without a CUDA device the context cannot be created: there is no CPU implementation behind the ABI."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    import golden_lib as G
    from gapprober_b200 import FigbirdContext, FigbirdError
    with pytest.raises(FigbirdError) as e:
        FigbirdContext(G.Case("lib180_100bp_60x").figbird_model())
    assert e.value.status == -3     # FB_ECUDA


def test_unknown_mode_is_refused(built):
    """f may be 0 or 1 (figbird.cpp:2578, 1814-1815); anything else is refused before CUDA is touched."""
    import golden_lib as G
    from gapprober_b200 import FigbirdContext, FigbirdError
    m = G.Case("logmode_f1").figbird_model()
    m.f_mode = 2
    with pytest.raises(FigbirdError) as e:
        FigbirdContext(m)
    assert e.value.status == -6


def test_dropin_binary_fails_loudly_without_gpu(built, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    import gzip
    import shutil
    from gapprober_b200 import build
    src = os.path.join(ROOT, "tests", "golden", "preamble")
    wd = str(tmp_path)
    os.makedirs(os.path.join(wd, "Gaps"))
    for f in ("myout.sam", "scaffolds.fa"):
        open(os.path.join(wd, f), "wb").write(gzip.open(os.path.join(src, f + ".gz")).read())
    for f in ("stat.txt", "gapInfo.txt"):
        shutil.copy(os.path.join(src, f), wd)
    for k in range(2):
        open(os.path.join(wd, "Gaps", f"gaps_{k}.sam"), "w").close()
    r = subprocess.run([build.build_host(), "scaffolds.fa", " 0", "250"], cwd=wd, stdout=subprocess.PIPE, text=True)
    assert r.returncode == 1
    assert "fb_create failed" in r.stdout or "cannot load" in r.stdout
    assert not os.path.exists(os.path.join(wd, "gapout.txt"))
