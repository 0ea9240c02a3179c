"""ctypes access to oracle/liboracle.so (the CPU restatement) and readers for the reference taps.

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
"""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")


class OrcModel(C.Structure):
    _fields_ = [
        ("max_read_len", C.c_int),
        ("error_pos", C.POINTER(C.c_double)),
        ("in_pos", C.POINTER(C.c_double)),
        ("del_pos", C.POINTER(C.c_double)),
        ("error_type", (C.c_double * 5) * 5),
        ("ild", C.POINTER(C.c_double)),
        ("ild_len", C.c_int),
        ("thr_min", C.c_int),
        ("thr_max", C.c_int),
        ("gap_prob_cutoff", C.c_int),
        ("max_distance", C.c_int),
        ("num_iter", C.c_int),
        ("f_mode", C.c_int),
    ]


class OrcRecord(C.Structure):
    _fields_ = [("pos1", C.c_int), ("len1", C.c_int), ("strand1", C.c_int), ("seq2", C.c_char_p)]


class OrcResult(C.Structure):
    _fields_ = [
        ("gap_min", C.c_int),
        ("gap_max", C.c_int),
        ("best_estimate", C.c_int),
        ("best_lik", C.c_double),
        ("total_reads", C.c_int),
        ("discarded", C.c_int),
        ("filled_len", C.c_int),
        ("filled", C.c_char_p),
        ("lik_per_estimate", C.POINTER(C.c_double)),
        ("dump_estimate", C.c_int),
        ("probs_dump", C.POINTER(C.c_double)),
        ("counts_dump", C.POINTER(C.c_double)),
        ("best_string", C.c_char_p),
    ]


_lib = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "all"])


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so):
            build_oracle()
        _lib = C.CDLL(so)
        _lib.orc_fill_gap.restype = C.c_int
        _lib.orc_fill_gap.argtypes = [C.POINTER(OrcModel), C.c_char_p, C.c_long, C.c_long, C.c_int,
                                      C.POINTER(OrcRecord), C.c_int, C.POINTER(OrcResult)]
        _lib.orc_fill_gap_sharded.restype = C.c_int
        _lib.orc_fill_gap_sharded.argtypes = [C.POINTER(OrcModel), C.c_char_p, C.c_long, C.c_long, C.c_int,
                                              C.POINTER(OrcRecord), C.c_int, C.c_int, C.POINTER(OrcResult)]
        _lib.orc_fill_gap_subset.restype = C.c_int
        _lib.orc_fill_gap_subset.argtypes = [C.POINTER(OrcModel), C.c_char_p, C.c_long, C.c_long, C.c_int,
                                             C.POINTER(OrcRecord), C.c_int, C.c_int, C.c_char_p, C.POINTER(OrcResult)]
        _lib.orc_gap_range.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib.orc_count_likelihoods.restype = C.c_long
        _lib.orc_count_likelihoods.argtypes = [C.POINTER(OrcModel), C.c_long, C.c_int, C.POINTER(OrcRecord), C.c_int]
    return _lib


class Model:
    """The hot path's constant inputs (what the reference's preamble computes)."""

    def __init__(self, max_read_len, ild, thr_min, thr_max, cutoff, error_type, error_pos, in_pos, del_pos,
                 max_distance=200, num_iter=5, f_mode=0):
        self.max_read_len = int(max_read_len)
        self.ild = np.ascontiguousarray(ild, dtype=np.float64)
        self.thr_min, self.thr_max, self.cutoff = int(thr_min), int(thr_max), int(cutoff)
        self.error_type = np.ascontiguousarray(error_type, dtype=np.float64).reshape(5, 5)
        self.error_pos = np.ascontiguousarray(error_pos, dtype=np.float64)
        self.in_pos = np.ascontiguousarray(in_pos, dtype=np.float64)
        self.del_pos = np.ascontiguousarray(del_pos, dtype=np.float64)
        self.max_distance, self.num_iter, self.f_mode = int(max_distance), int(num_iter), int(f_mode)

    @staticmethod
    def load(path):
        """Read fb_model.bin (format in oracle/patch_ref.py)."""
        raw = open(path, "rb").read()
        assert raw[:4] == b"FBM1", "bad model file"
        mrl, mis, tmin, tmax, cut, D = struct.unpack_from("<6i", raw, 4)
        off = 4 + 24
        d = np.frombuffer(raw, dtype="<f8", offset=off)
        et = d[:25].reshape(5, 5)
        ep, ip, dp = d[25:25 + mrl], d[25 + mrl:25 + 2 * mrl], d[25 + 2 * mrl:25 + 3 * mrl]
        ild = d[25 + 3 * mrl:25 + 3 * mrl + mis]
        assert ild.size == mis
        return Model(mrl, ild, tmin, tmax, cut, et, ep, ip, dp, max_distance=D)

    def as_orc(self):
        m = OrcModel()
        m.max_read_len = self.max_read_len
        dp = C.POINTER(C.c_double)
        m.error_pos = self.error_pos.ctypes.data_as(dp)
        m.in_pos = self.in_pos.ctypes.data_as(dp)
        m.del_pos = self.del_pos.ctypes.data_as(dp)
        for i in range(5):
            for j in range(5):
                m.error_type[i][j] = self.error_type[i, j]
        m.ild = self.ild.ctypes.data_as(dp)
        m.ild_len = self.ild.size
        m.thr_min, m.thr_max, m.gap_prob_cutoff = self.thr_min, self.thr_max, self.cutoff
        m.max_distance, m.num_iter, m.f_mode = self.max_distance, self.num_iter, self.f_mode
        return m


def read_gap_file(path):
    """(pos1, len1, strand1, seq2) per line pair, as placeReads parses it (figbird.cpp:1718-1801)."""
    recs = []
    with open(path, "rb") as f:
        lines = [ln for ln in f.read().split(b"\n") if ln and not ln.startswith(b"@")]
    for a, b in zip(lines[0::2], lines[1::2]):
        fa, fb = a.split(b"\t"), b.split(b"\t")
        recs.append((int(fa[3]), len(fa[6]), (int(fa[1]) & 16) >> 4, fb[6]))
    return recs


def gap_range(orig):
    a, b = C.c_int(), C.c_int()
    lib().orc_gap_range(orig, C.byref(a), C.byref(b))
    return a.value, b.value


def fill_gap(model, contig, gap_start, orig_gap, recs, dump_estimate=0, shards=1, estimates=None):
    """Run the oracle on one gap.  contig: bytes (upper-cased scaffold).  Returns a dict.
    shards > 1: the record-sharded grouping of the M-step sums (orc_fill_gap_sharded) instead of the reference's.
    estimates: scan only these estimates (orc_fill_gap_subset); the others read NaN in 'liks'."""
    L = lib()
    m = model.as_orc()
    arr = (OrcRecord * max(1, len(recs)))()
    keep = []
    for i, (p, l1, s1, seq) in enumerate(recs):
        arr[i].pos1, arr[i].len1, arr[i].strand1 = p, l1, s1
        keep.append(C.c_char_p(seq))
        arr[i].seq2 = keep[-1]
    gmin, gmax = gap_range(orig_gap)
    ne = max(0, gmax - gmin + 1)
    cap = max(gmax, 2 * orig_gap) + 2
    res = OrcResult()
    filled = C.create_string_buffer(cap)
    best = C.create_string_buffer(cap)
    liks = np.zeros(max(1, ne))
    res.filled = C.cast(filled, C.c_char_p)
    res.best_string = C.cast(best, C.c_char_p)
    res.lik_per_estimate = liks.ctypes.data_as(C.POINTER(C.c_double))
    probs = counts = None
    res.dump_estimate = int(dump_estimate)
    if dump_estimate > 0:
        n = (dump_estimate + 2 * model.max_distance) * 5
        probs, counts = np.zeros(n), np.zeros(n)
        res.probs_dump = probs.ctypes.data_as(C.POINTER(C.c_double))
        res.counts_dump = counts.ctypes.data_as(C.POINTER(C.c_double))
    scan = None
    if estimates is not None:
        mask = bytearray(max(1, ne))
        for e in estimates:
            mask[int(e) - gmin] = 1
        scan = bytes(mask)
    rc = L.orc_fill_gap_subset(C.byref(m), contig, len(contig), gap_start, orig_gap, arr, len(recs), int(shards), scan, C.byref(res))
    if rc == -3 and estimates is not None:
        # finalize counted >= 10000 reads for base A at some position: the reference's checkDuplicate then indexes past
        # its stack array (figbird.cpp:1967-1982) — undefined there, so only the scan's outputs are returned
        return dict(undefined_finalize=True, best_string=best.value, liks=liks[:ne].copy(), filled=None)
    if rc != 0:
        raise RuntimeError(f"orc_fill_gap failed: {rc}")
    return dict(gap_min=res.gap_min, gap_max=res.gap_max, best_estimate=res.best_estimate, best_lik=res.best_lik,
                total_reads=res.total_reads, discarded=res.discarded, filled=filled.value[:res.filled_len],
                best_string=best.value, liks=liks[:ne].copy(),
                probs=None if probs is None else probs.reshape(-1, 5),
                counts=None if counts is None else counts.reshape(-1, 5))


def count_likelihoods(model, gap_start, gp, recs):
    m = model.as_orc()
    arr = (OrcRecord * max(1, len(recs)))()
    keep = []
    for i, (p, l1, s1, seq) in enumerate(recs):
        arr[i].pos1, arr[i].len1, arr[i].strand1 = p, l1, s1
        keep.append(C.c_char_p(seq))
        arr[i].seq2 = keep[-1]
    return lib().orc_count_likelihoods(C.byref(m), gap_start, gp, arr, len(recs))


def read_fasta(path):
    """Scaffolds as the reference loads them: names up to the first blank, sequences upper-cased."""
    names, seqs, cur = [], [], []
    with open(path, "rb") as f:
        for ln in f:
            if ln.startswith(b";"):
                continue
            if ln.startswith(b">"):
                if cur:
                    seqs.append(b"".join(cur).upper())
                    cur = []
                names.append(ln[1:].split()[0])
            else:
                cur.append(ln.rstrip(b"\n"))
    seqs.append(b"".join(cur).upper())
    return names, seqs


def read_gap_info(path):
    out = []
    with open(path) as f:
        for ln in f:
            if ln.strip():
                c, s, l = ln.split()[:3]
                out.append((int(c), int(s), int(l)))
    return out


def parse_dump(path):
    """fb_dump.txt -> {gapfile: {"est": {estimate: lik}, "rows": {estimate: ndarray[rows,10]}, "best": (...)}}"""
    out = {}
    cur_rows, cur_key = None, None
    with open(path) as f:
        for ln in f:
            p = ln.rstrip("\n").split("\t")
            if p[0] == "EST":
                g = out.setdefault(p[1], {"est": {}, "rows": {}, "best": None})
                g["est"][int(p[2])] = float.fromhex(p[3])
                cur_rows = []
                g["rows"][int(p[2])] = cur_rows
            elif p[0] == "P":
                cur_rows.append([float.fromhex(x) for x in p[2:12]])
            elif p[0] == "BEST":
                g = out.setdefault(p[1], {"est": {}, "rows": {}, "best": None})
                g["best"] = (int(p[2]), float.fromhex(p[3]), p[4] if len(p) > 4 else "")
    for g in out.values():
        for k in list(g["rows"]):
            g["rows"][k] = np.array(g["rows"][k], dtype=np.float64) if g["rows"][k] else None
    return out
