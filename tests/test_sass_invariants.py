"""Build-level invariants of the EM kernels, read off the SASS of the in-tree library (cuobjdump; no GPU needed).

What the parity suite proves on a GPU rests on properties of the generated code that a compiler flag or a careless
edit can silently break; they are checked here on every CPU run:
  * no fused multiply-add in the E-step's product chain: the reference evaluates P*(1-e) + e*Q as two rounded products
    and a rounded sum and multiplies the running score in read order (figbird.cpp:1855, 1862) - a DFMA in the hot loops
    would change bits (DFMA legitimately appears elsewhere: the expansions of __ddiv_rn and the log/pow restatements);
  * the chunk staging really is a bulk asynchronous copy (UBLKCP) in every EM variant;
  * the E loop's shape: per trip of 8 bases a two-chain loop with 16 LDS.128, 48 DMUL, 16 DADD;
  * no local-memory traffic (spills) inside the hot loops;
  * the task loop of the main variant fits the 32 KB instruction-cache level (growing past it
    cost 5-6 %).
"""
import os
import re
import shutil
import subprocess
from collections import Counter

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gapprober_b200", "libfigbird_cuda.so")


def _cuobjdump():
    for cand in (shutil.which("cuobjdump"), "/usr/local/cuda/bin/cuobjdump"):
        if cand and os.path.exists(cand):
            return cand
    return None


@pytest.fixture(scope="module")
def kernels():
    """{mangled kernel name: [(address, instruction text)]} of the library's sm_100a code."""
    exe = _cuobjdump()
    if exe is None:
        pytest.skip("cuobjdump not found")
    if not os.path.exists(LIB):
        from gapprober_b200 import build
        build.build_library()
    out = subprocess.run([exe, "-sass", LIB], stdout=subprocess.PIPE, check=True).stdout.decode()
    fns, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = fns.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    assert "arch = sm_100a" in out or "sm_100a" in out
    return fns


def _op(text):
    return re.sub(r"^@!?U?P\d+\s+", "", text).split()[0]


def _loops(ins):
    """Backward branches: [(first index, last index)] of every loop body."""
    at = {a: i for i, (a, _) in enumerate(ins)}
    out = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in at:
                out.append((at[tgt], i))
    return out


def _em_variants(kernels, f1=None):
    for name, ins in kernels.items():
        m = re.search(r"fb_em_kernelILb(\d)ELb(\d)ELi(\d+)ELb(\d)E", name)
        if m and (f1 is None or int(m.group(4)) == int(f1)):
            yield name, ins, dict(gt=int(m.group(1)), clu=int(m.group(2)), nt=int(m.group(3)), f1=int(m.group(4)))


def test_every_em_variant_is_built(kernels):
    got = sorted((v["gt"], v["clu"], v["nt"], v["f1"]) for _, _, v in _em_variants(kernels))
    want = sorted((gt, clu, nt, f1) for gt, clu in ((0, 0), (1, 0), (0, 1)) for nt in (512, 1024) for f1 in (0, 1))
    assert got == want


def test_chunk_staging_is_a_bulk_async_copy(kernels):
    for name, ins, v in _em_variants(kernels):
        ops = Counter(_op(t) for _, t in ins)
        assert sum(n for o, n in ops.items() if o.startswith("UBLKCP")) >= 5, (v, "the five bulk copies per staged chunk are gone")
        assert any(o.startswith("SYNCS") for o in ops), (v, "no mbarrier operations")


def test_e_step_hot_loops_have_no_fma_and_no_spills(kernels):
    """The shared-memory variants in f=0: every loop that is dominated by {P,Q} loads is DMUL/DADD only."""
    for name, ins, v in _em_variants(kernels, f1=0):
        if v["gt"]:
            continue                      # the global-table variant reads {P,Q} with LDG: covered by the parity tests
        hot = []
        for s, e in _loops(ins):
            body = [t for _, t in ins[s:e + 1]]
            ops = Counter(_op(t) for t in body)
            if ops["LDS.128"] >= 4 and len(body) < 220 and ops["DMUL"] >= 3 * ops["LDS.128"] - 2:
                hot.append((len(body), ops))
                assert ops["DFMA"] == 0, (v, "fused multiply-add in an E-step loop")
                assert ops["LDL"] == 0 and ops["STL"] == 0, (v, "spill inside an E-step loop")
                assert ops["DADD"] == ops["LDS.128"], (v, "one rounded sum per {P,Q} pair")
        assert hot, (v, "E-step loops not found")
        # the main loop: two chains, eight bases per trip
        assert any(o["LDS.128"] == 16 and o["DMUL"] == 48 and o["DADD"] == 16 for _, o in hot), (v, [h[0] for h in hot])


def test_task_loop_fits_the_instruction_cache_level(kernels):
    """Main variant <GT=0, CLU=0, 512 threads, f=0>: the smallest loop holding the E loops and the M-step's score loads
    (= the task loop of a round) is at most 32 KB of code."""
    (name, ins, v), = [k for k in _em_variants(kernels, f1=0) if k[2]["gt"] == 0 and k[2]["clu"] == 0 and k[2]["nt"] == 512]
    loops = _loops(ins)
    stats = []
    for s, e in loops:
        ops = Counter(_op(t) for _, t in ins[s:e + 1])
        stats.append((s, e, ops))
    e_loops = [(s, e) for s, e, o in stats if e - s < 220 and o["LDS.128"] >= 8]
    m_loops = [(s, e) for s, e, o in stats if e - s < 120 and o["LDS.64"] >= 8 and o["LDS.128"] == 0]
    assert e_loops and m_loops
    lo = min(s for s, _ in e_loops + m_loops)
    hi = max(e for _, e in e_loops + m_loops)
    enclosing = [(e - s + 1) for s, e, _ in stats if s <= lo and e >= hi]
    assert enclosing, "no loop encloses the E and M loops"
    task_loop = min(enclosing)
    assert task_loop * 16 <= 32 * 1024, f"task loop is {task_loop} instructions = {task_loop * 16} bytes"
    assert len(ins) < 5200, f"main EM variant has grown to {len(ins)} instructions"


def test_register_and_parameter_budgets():
    """Two 512-thread CTAs (or one of 1024 threads) per SM need at most 64 registers per thread; the kernel parameters
    (both copies of the {1-e, e} table and the first-iteration terms travel in them) must stay below the 32 KB limit."""
    exe = _cuobjdump()
    if exe is None or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the library not found")
    out = subprocess.run([exe, "-res-usage", LIB], stdout=subprocess.PIPE, check=True).stdout.decode().splitlines()
    seen = 0
    for i, ln in enumerate(out):
        m = re.search(r"Function .*fb_em_kernelILb(\d)ELb(\d)ELi(\d+)ELb(\d)E", ln)
        if not m:
            continue
        gt, clu, nt, f1 = (int(x) for x in m.groups())
        use = dict((k, int(v)) for k, v in re.findall(r"(\w+(?:\[0\])?):(\d+)", out[i + 1]))
        seen += 1
        if not gt or nt == 1024:
            assert use["REG"] <= 64, (gt, clu, nt, f1, use)
        assert use["STACK"] <= 256, (gt, clu, nt, f1, use)           # a few spilled words outside the hot loops, no more
        assert use["CONSTANT[0]"] < 32 * 1024, (gt, clu, nt, f1, use)
        assert use["SHARED"] - 1024 <= 1024, (gt, clu, nt, f1, use)  # static shared memory (the planner's budget assumes < 1 KB)
    assert seen == 12
