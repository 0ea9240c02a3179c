"""The oracle (C restatement, oracle/figbird_oracle.c) pinned against the reference's own outputs: every golden
fixture was produced by running the unmodified reference splitter and an instrumented reference figbird
(tests/golden/make_golden.py).  Bit-exact on everything: integer/byte work and FP64 alike."""
import numpy as np
import pytest

import golden_lib as G
import oracle_lib as O


@pytest.mark.parametrize("name", G.ALL_CASES)
def test_oracle_matches_reference_golden(name, built):
    case = G.Case(name)
    m = case.oracle_model()
    for gi, g in enumerate(case.gaps):
        r = O.fill_gap(m, g["window"], case.D, g["orig"], g["recs"], dump_estimate=g["best_est"])
        assert np.array_equal(r["liks"], g["liks"]), f"{name} gap {gi}: per-estimate likelihoods"
        assert r["best_estimate"] == g["best_est"] and r["best_lik"] == g["best_lik"]
        assert r["best_string"] == g["best_string"]
        assert r["filled"] == g["filled"]
        assert (r["total_reads"], r["discarded"]) == g["counts"]
        assert np.array_equal(r["probs"], g["rows"][:, :5], equal_nan=True)
        assert np.array_equal(r["counts"], g["rows"][:, 5:], equal_nan=True)


def test_golden_covers_the_hard_cases():
    """The fixtures must contain the behaviours SURVEY §0 calls load-bearing: NaN rows, likelihood plateaus
    (exact ties resolved by '>=' last-wins), reads with N, multi-hit duplicates, both f modes."""
    nan_rows = ties = n_reads = 0
    for name in G.ALL_CASES:
        case = G.Case(name)
        for g in case.gaps:
            nan_rows += int(np.isnan(g["rows"][:, 0]).sum())
            liks = g["liks"]
            ties += int((liks == liks.max()).sum() > 1)
            n_reads += sum(b"N" in s for _, _, _, s in g["recs"])
    assert nan_rows > 0 and ties > 0 and n_reads > 0


def test_gap_range_float_quirks(built):
    """fillGap's range uses float fractions (figbird.cpp:2438-2453): e.g. 400*1.05f truncates to 419, not 420."""
    assert O.gap_range(400) == (380, 419)
    assert O.gap_range(99) == (1, 148)
    assert O.gap_range(100) == (85, 150)
    assert O.gap_range(500) == (475, 525)
    assert O.gap_range(1) == (1, 1)


def test_oracle_refuses_undefined_territory(built):
    case = G.Case("lib180_100bp_60x")
    g = case.gaps[0]
    with pytest.raises(RuntimeError):
        O.fill_gap(case.oracle_model(), g["window"][50:], case.D - 50, g["orig"], g["recs"])   # gap closer than D to the scaffold start


def test_estimate_subset_scan_equals_the_full_scan(built):
    """orc_fill_gap_subset (the test aid for gaps too big to scan fully on one core): a scanned estimate has the bits
    it has in the full scan, the others read NaN, and with the full scan's winner in the subset the fill is the same."""
    import golden_lib as G
    case = G.Case("multi_gap_err1pct")
    for g in case.gaps[:3]:
        full = O.fill_gap(case.oracle_model(), g["window"], case.D, g["orig"], g["recs"])
        sub = sorted({full["gap_min"], full["best_estimate"], full["gap_max"]})
        part = O.fill_gap(case.oracle_model(), g["window"], case.D, g["orig"], g["recs"], estimates=sub)
        for e in range(full["gap_min"], full["gap_max"] + 1):
            v = part["liks"][e - full["gap_min"]]
            assert (v == full["liks"][e - full["gap_min"]]) if e in sub else np.isnan(v)
        assert (part["best_estimate"], part["filled"], part["discarded"]) == (full["best_estimate"], full["filled"], full["discarded"])
