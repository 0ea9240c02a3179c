"""Host-side mirror of the reference's figbird interface over the C ABI (include/figbird_cuda.h).

The reference exposes `class GapFiller` (figbird.cpp:1517-2532) driven by main()'s gap loop
(figbird.cpp:2819-2855) and talks to the rest of GapProber through files in the working directory.
This module keeps those names and meanings:

    model = FigbirdModel.load("fb_model.bin")            # what the preamble leaves in globals
    with FigbirdContext(model) as fb:
        gf = fb.GapFiller()                               # allocate + initGapFiller
        gf.initGapFiller(contigNo, gapStart, originalGap, "Gaps/gaps_0.sam", contig)
        fb.fillGaps()                                     # fillGap() of every registered gap, on the GPU
        n, s = gf.getConcensus()                          # figbird.cpp:2050-2063

Everything numeric happens in libfigbird_cuda.so (hand-written sm_100a kernels).  There is no CPU
fallback: importing works anywhere (so the ABI can be inspected), creating a context needs a GPU.
"""
import ctypes as C
import os
import struct

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfigbird_cuda.so")

FB_OK = 0


class FigbirdError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"figbird_cuda status {status}: {msg}")
        self.status = status


class _FbModel(C.Structure):
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


class _FbRecord(C.Structure):
    _fields_ = [("pos1", C.c_int), ("len1", C.c_int), ("strand1", C.c_int), ("seq2", C.c_char_p), ("len2", C.c_int)]


class _FbGapDesc(C.Structure):
    _fields_ = [("contig_no", C.c_int), ("gap_start", C.c_long), ("orig_gap", C.c_int), ("contig", C.c_char_p),
                ("contig_len", C.c_long), ("gap_file", C.c_char_p)]


class _FbGapResult(C.Structure):
    _fields_ = [
        ("gap_min", C.c_int), ("gap_max", C.c_int), ("best_estimate", C.c_int), ("best_lik", C.c_double),
        ("total_reads", C.c_int), ("discarded", C.c_int), ("filled_len", C.c_int), ("likelihoods", C.c_long),
    ]


class FbStats(C.Structure):
    _fields_ = [
        ("n_gaps", C.c_long), ("n_units", C.c_long), ("n_records", C.c_long),
        ("likelihoods", C.c_long), ("finalize_likelihoods", C.c_long), ("base_terms", C.c_long),
        ("em_kernel_ms", C.c_double), ("finalize_kernel_ms", C.c_double), ("step_ms", C.c_double),
        ("em_launches", C.c_int), ("finalize_launches", C.c_int),
        ("h2d_bytes", C.c_long), ("d2h_bytes", C.c_long), ("algorithmic_bytes", C.c_long),
        ("aux_launches", C.c_int),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# Every symbol include/figbird_cuda.h declares; tests check the library exports all of them.
ABI_SYMBOLS = [
    "fb_create", "fb_warmup", "fb_destroy", "fb_add_gap", "fb_add_gap_file", "fb_add_gap_files", "fb_set_partition", "fb_set_partition_units", "fb_run_scan", "fb_run_finalize", "fb_scan_liks", "fb_merge_liks", "fb_owns_gap", "fb_partition_plan", "fb_gap_cost", "fb_run", "fb_upload",
    "fb_run_device", "fb_fill_gap_files", "fb_result", "fb_estimate_liks", "fb_best_string", "fb_dump_profile", "fb_get_stats",
    "fb_clear_gaps", "fb_gap_range", "fb_last_error", "fb_status_str", "fb_abi_version",
    "fb_set_partition_records", "fb_shard_sizes", "fb_shard_buffers", "fb_shard_scan_iter", "fb_shard_scan_finish", "fb_shard_scan_all",
    "fb_shard_counts", "fb_shard_reset", "fb_measure_smem_peak", "fb_devmath_eval", "fb_ipc_export", "fb_ipc_open", "fb_ipc_close",
]

_lib = None


def load_library():
    """dlopen libfigbird_cuda.so and declare the prototypes.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("FIGBIRD_CUDA_LIB", LIB_PATH)     # same override the drop-in binary honours
    if not os.path.exists(path):
        raise FigbirdError(-3, f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.fb_create.argtypes = [C.POINTER(_FbModel), C.c_int, C.POINTER(vp)]
    L.fb_warmup.argtypes = [C.c_int]
    L.fb_destroy.argtypes = [vp]
    L.fb_destroy.restype = None
    L.fb_add_gap.argtypes = [vp, C.c_int, C.c_long, C.c_int, C.c_char_p, C.c_long, C.POINTER(_FbRecord), C.c_int]
    L.fb_add_gap_file.argtypes = [vp, C.c_int, C.c_long, C.c_int, C.c_char_p, C.c_long, C.c_char_p]
    L.fb_add_gap_files.argtypes = [vp, C.POINTER(_FbGapDesc), C.c_int, C.POINTER(C.c_int), C.c_int]
    L.fb_fill_gap_files.argtypes = [vp, C.POINTER(_FbGapDesc), C.c_int, C.POINTER(C.c_int), C.c_int]
    L.fb_set_partition.argtypes = [vp, C.c_int, C.c_int]
    L.fb_set_partition_units.argtypes = [vp, C.c_int, C.c_int]
    L.fb_run_scan.argtypes = [vp]
    L.fb_run_finalize.argtypes = [vp]
    L.fb_scan_liks.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.c_int]
    L.fb_merge_liks.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.c_int]
    L.fb_owns_gap.argtypes = [vp, C.c_int]
    L.fb_partition_plan.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.fb_gap_cost.argtypes = [vp, C.c_int]
    L.fb_gap_cost.restype = C.c_double
    L.fb_run.argtypes = [vp]
    L.fb_upload.argtypes = [vp]
    L.fb_run_device.argtypes = [vp]
    L.fb_result.argtypes = [vp, C.c_int, C.POINTER(_FbGapResult), C.c_char_p, C.c_int]
    L.fb_estimate_liks.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.c_int]
    L.fb_best_string.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
    L.fb_dump_profile.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.fb_get_stats.argtypes = [vp, C.POINTER(FbStats)]
    L.fb_clear_gaps.argtypes = [vp]
    L.fb_gap_range.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.fb_gap_range.restype = None
    L.fb_last_error.argtypes = [vp]
    L.fb_last_error.restype = C.c_char_p
    L.fb_status_str.argtypes = [C.c_int]
    L.fb_status_str.restype = C.c_char_p
    L.fb_abi_version.restype = C.c_int
    dpp = C.POINTER(C.c_void_p)
    L.fb_set_partition_records.argtypes = [vp, C.c_int, C.c_int]
    L.fb_shard_sizes.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.fb_shard_buffers.argtypes = [vp, dpp, dpp, dpp, dpp]
    L.fb_shard_scan_all.argtypes = [vp, dpp, dpp, dpp, dpp]
    L.fb_shard_scan_iter.argtypes = [vp, C.c_int, dpp, vp]
    L.fb_shard_scan_finish.argtypes = [vp, dpp, dpp]
    L.fb_shard_counts.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.fb_shard_reset.argtypes = [vp]
    L.fb_measure_smem_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.fb_devmath_eval.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_long]
    L.fb_ipc_export.argtypes = [vp, vp, C.c_char_p]
    L.fb_ipc_open.argtypes = [vp, C.c_char_p, dpp]
    L.fb_ipc_close.argtypes = [vp, vp]
    _lib = L
    return L


def partition_plan(costs, world):
    """Owner rank of every work item: descending cost, each to the least loaded rank (the rule fb_set_partition uses)."""
    cost = np.ascontiguousarray(costs, dtype=np.float64)
    owner = np.zeros(cost.size, dtype=np.int32)
    rc = load_library().fb_partition_plan(cost.ctypes.data_as(C.POINTER(C.c_double)), cost.size, int(world),
                                          owner.ctypes.data_as(C.POINTER(C.c_int)))
    if rc != FB_OK:
        raise FigbirdError(rc, "fb_partition_plan")
    return owner


def measure_smem_peak(device=0):
    """Measured shared-memory read bandwidth of the device in GB/s (conflict-free LDS.128 on every SM)."""
    v = C.c_double()
    rc = load_library().fb_measure_smem_peak(int(device), C.byref(v))
    if rc != FB_OK:
        raise FigbirdError(rc, "fb_measure_smem_peak")
    return v.value


def device_math(what, x, device=0):
    """The device's restatement of glibc's log ("log") or pow(10, .) ("pow10") applied to an array (parity tap)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    dp = C.POINTER(C.c_double)
    rc = load_library().fb_devmath_eval(int(device), {"log": 0, "pow10": 1}[what], x.ctypes.data_as(dp), out.ctypes.data_as(dp), x.size)
    if rc != FB_OK:
        raise FigbirdError(rc, "fb_devmath_eval")
    return out


def gap_range(original_gap):
    """Estimate range scanned by fillGap (figbird.cpp:2438-2453)."""
    a, b = C.c_int(), C.c_int()
    load_library().fb_gap_range(int(original_gap), C.byref(a), C.byref(b))
    return a.value, b.value


class FigbirdModel:
    """Constant inputs of the hot path: the globals the reference's preamble fills (figbird.cpp:475-823, 2778-2794)."""

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
        """Read a model file ("FBM1": 6 int32 + doubles; written by the host preamble or by the instrumented reference)."""
        raw = open(path, "rb").read()
        if raw[:4] != b"FBM1":
            raise ValueError("not a figbird model file")
        mrl, mis, tmin, tmax, cut, D = struct.unpack_from("<6i", raw, 4)
        d = np.frombuffer(raw, dtype="<f8", offset=28)
        et = d[:25].reshape(5, 5)
        ep, ip, dp = d[25:25 + mrl], d[25 + mrl:25 + 2 * mrl], d[25 + 2 * mrl:25 + 3 * mrl]
        ild = d[25 + 3 * mrl:25 + 3 * mrl + mis]
        return FigbirdModel(mrl, ild, tmin, tmax, cut, et, ep, ip, dp, max_distance=D)

    def save(self, path):
        with open(path, "wb") as f:
            f.write(b"FBM1")
            f.write(struct.pack("<6i", self.max_read_len, self.ild.size, self.thr_min, self.thr_max, self.cutoff, self.max_distance))
            for a in (self.error_type, self.error_pos, self.in_pos, self.del_pos, self.ild):
                f.write(np.ascontiguousarray(a, dtype="<f8").tobytes())

    def _as_c(self):
        m = _FbModel()
        dp = C.POINTER(C.c_double)
        m.max_read_len = self.max_read_len
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
    """Line pairs of Gaps/gaps_<n>.sam as placeReads parses them (figbird.cpp:1718-1801):
    (pos1, len1, strand1, seq2) per pair; '@' lines skipped, a trailing unpaired line dropped."""
    recs = []
    with open(path, "rb") as f:
        lines = [ln for ln in f.read().split(b"\n") if ln and not ln.startswith(b"@")]
    for a, b in zip(lines[0::2], lines[1::2]):
        fa, fb = a.split(b"\t"), b.split(b"\t")
        recs.append((int(fa[3]), len(fa[6]), (int(fa[1]) & 16) >> 4, fb[6]))
    return recs


def read_scaffolds(path):
    """Scaffold FASTA as figbird's main() loads it (figbird.cpp:2586-2658): names up to the first blank, upper-cased bases."""
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
    """gapInfo.txt rows (contigNo, gapStart, gapLength), main.cpp:965."""
    out = []
    with open(path) as f:
        for ln in f:
            if ln.strip():
                c, s, l = ln.split()[:3]
                out.append((int(c), int(s), int(l)))
    return out


class GapFiller:
    """One gap.  Method names follow the reference class (figbird.cpp:1517-2532)."""

    def __init__(self, ctx):
        self._ctx = ctx
        self.index = None
        self.originalGap = None
        self.gapStart = None
        self.contigNo = None

    def initGapFiller(self, contigNo, gapStart, originalGap, gapFile, contig):
        """figbird.cpp:1610-1618 + the gap-file parse.  gapFile: path (parsed inside the library, as placeReads
        does), or an already parsed list of (pos1, len1, strand1, seq2)."""
        self.contigNo, self.gapStart, self.originalGap = int(contigNo), int(gapStart), int(originalGap)
        if isinstance(gapFile, (str, bytes, os.PathLike)):
            self.index = self._ctx._add_gap_file(self.contigNo, self.gapStart, self.originalGap, contig, os.fsencode(gapFile))
        else:
            self.index = self._ctx._add_gap(self.contigNo, self.gapStart, self.originalGap, contig, list(gapFile))
        return self

    def _result(self):
        return self._ctx._result(self.index)

    def getConcensus(self):
        """(length, string) — figbird.cpp:2050-2063."""
        r = self._result()
        return r["filled_len"], r["filled"]

    @property
    def maxGapEstimate(self):
        return self._result()["best_estimate"]

    @property
    def maxLikelihood(self):
        return self._result()["best_lik"]

    def result(self):
        return self._result()

    def estimateLikelihoods(self):
        return self._ctx._liks(self.index)

    def bestString(self):
        return self._ctx._best_string(self.index)

    def dumpProfile(self, estimate):
        """(probsGap, countsGap) after the last EM iteration of one estimate, [(estimate+2D), 5] each."""
        return self._ctx._dump(self.index, estimate)


class GapBatch:
    """The descriptors of many gaps ([(contigNo, gapStart, originalGap, gapFilePath, contig bytes)]) laid out once for
    fb_add_gap_files, for callers that register the same gap set repeatedly."""

    def __init__(self, gaps):
        n = len(gaps)
        self.n = n
        self.descs = (_FbGapDesc * max(1, n))()
        self.handles = (C.c_int * max(1, n))()
        self._keep = []
        self.meta = []
        self._fillers = None
        for i, (c, s, l, path, contig) in enumerate(gaps):
            p = os.fsencode(path)
            self._keep.append((p, contig))
            d = self.descs[i]
            d.contig_no, d.gap_start, d.orig_gap = int(c), int(s), int(l)
            d.contig, d.contig_len, d.gap_file = contig, len(contig), p
            self.meta.append((int(c), int(s), int(l)))


class FigbirdContext:
    """A model bound to one GPU; owns the registered gaps.  fillGaps() = GapFiller::fillGap for all of them."""

    def __init__(self, model, device=0):
        self._L = load_library()
        self.model = model
        self._h = C.c_void_p()
        self._cm = model._as_c()
        rc = self._L.fb_create(C.byref(self._cm), int(device), C.byref(self._h))
        if rc != FB_OK:
            raise FigbirdError(rc, self._L.fb_status_str(rc).decode() + " (fb_create; is a CUDA device visible?)")
        self._gaps = []

    def close(self):
        if self._h:
            self._L.fb_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise FigbirdError(rc, self._L.fb_last_error(self._h).decode() or self._L.fb_status_str(rc).decode())
        return rc

    def GapFiller(self):
        g = GapFiller(self)
        self._gaps.append(g)
        return g

    def _add_gap(self, contig_no, gap_start, orig, contig, recs):
        arr = (_FbRecord * max(1, len(recs)))()
        for i, (p, l1, s1, seq) in enumerate(recs):
            arr[i].pos1, arr[i].len1, arr[i].strand1 = int(p), int(l1), int(s1)
            arr[i].seq2 = seq
            arr[i].len2 = len(seq)
        return self._check(self._L.fb_add_gap(self._h, contig_no, gap_start, orig, contig, len(contig), arr, len(recs)))

    def _add_gap_file(self, contig_no, gap_start, orig, contig, path):
        return self._check(self._L.fb_add_gap_file(self._h, contig_no, gap_start, orig, contig, len(contig), path))

    def add_gap_files(self, gaps, threads=0):
        """Register many gaps at once: gaps = [(contigNo, gapStart, originalGap, gapFilePath, contig bytes)] or a GapBatch
        prepared from such a list.  The library parses and packs the files on several host threads
        (fb_add_gap_files).  Returns the GapFiller objects."""
        return self._register(gaps, threads, self._L.fb_add_gap_files)

    def fill_gap_files(self, gaps, threads=0):
        """The whole gap loop in one call (fb_fill_gap_files): register + fill, with the host's parsing, packing and
        planning of most of the gap set overlapped with the device's work on the first batch.  Needs a context without
        gaps (clearGaps).  Same arguments and return value as add_gap_files; the results are ready on return."""
        return self._register(gaps, threads, self._L.fb_fill_gap_files)

    def _register(self, gaps, threads, call):
        batch = gaps if isinstance(gaps, GapBatch) else GapBatch(gaps)
        n = batch.n
        rc = call(self._h, batch.descs, n, batch.handles, int(threads))
        handles = np.frombuffer(batch.handles, dtype=np.int32, count=n)
        cached = batch._fillers
        if (rc == FB_OK and cached is not None and cached[0] is self and len(self._gaps) == (int(handles.min()) if n else 0)
                and np.array_equal(handles, cached[1])):
            out = cached[2]                      # same context, same handles as last time: the objects are still right
            self._gaps.extend(cached[3])         # ... in handle order
            return out
        out = []
        for (c, s, l), h in zip(batch.meta, handles.tolist()):
            if h < 0:
                continue
            g = GapFiller(self)
            g.contigNo, g.gapStart, g.originalGap, g.index = c, s, l, h
            while len(self._gaps) <= h:
                self._gaps.append(None)
            self._gaps[h] = g
            out.append(g)
        self._check(rc)
        if n and len(out) == n and len(self._gaps) == int(handles.min()) + n:   # every handle new and dense: cacheable
            base = int(handles.min())
            batch._fillers = (self, handles.copy(), out, self._gaps[base:base + n])
        return out

    def add_workdir(self, wd, scaffolds="scaffolds.fa", threads=0):
        """Register every gap of a GapProber working directory (gapInfo.txt + Gaps/), like figbird.cpp:2819-2842."""
        _, seqs = read_scaffolds(os.path.join(wd, scaffolds))
        gaps = [(c, s, l, os.path.join(wd, "Gaps", f"gaps_{gi}.sam"), seqs[c])
                for gi, (c, s, l) in enumerate(read_gap_info(os.path.join(wd, "gapInfo.txt")))]
        return self.add_gap_files(gaps, threads)

    def setPartition(self, rank, world):
        self._check(self._L.fb_set_partition(self._h, rank, world))

    def setPartitionUnits(self, rank, world):
        """Deal (gap, estimate) units instead of gaps (for gap sets dominated by few gaps)."""
        self._check(self._L.fb_set_partition_units(self._h, rank, world))

    def setPartitionRecords(self, rank, world):
        """Record-sharded scan: every rank scans every unit over its block of each gap's records (include/figbird_cuda.h)."""
        self._check(self._L.fb_set_partition_records(self._h, rank, world))
        self._world = world

    def shardSizes(self):
        """(doubles per partial-count buffer, doubles of the per-record maxima buffer), after upload()."""
        a, b = C.c_size_t(), C.c_size_t()
        self._check(self._L.fb_shard_sizes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def shardBuffers(self):
        """Device addresses of this context's own exchange buffers: (counts slot 0, counts slot 1, maxima, barrier flags)."""
        a, b, m, f = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self._L.fb_shard_buffers(self._h, C.byref(a), C.byref(b), C.byref(m), C.byref(f)))
        return a.value, b.value, m.value, f.value

    def shardScanAll(self, peer_counts0, peer_counts1, peer_maxprob, peer_flags):
        """The whole record-sharded scan in one call, ranks synchronised by flags in peer memory (one GPU per rank)."""
        self._check(self._L.fb_shard_scan_all(self._h, self._ptr_array(peer_counts0), self._ptr_array(peer_counts1),
                                              self._ptr_array(peer_maxprob), self._ptr_array(peer_flags)))

    @staticmethod
    def _ptr_array(ptrs):
        return (C.c_void_p * len(ptrs))(*[int(x) for x in ptrs])

    def shardScanIter(self, it, peer_counts, out_counts):
        """One EM iteration over the local record blocks; peer_counts: device addresses of every rank's previous partial counts."""
        arr = self._ptr_array(peer_counts) if peer_counts else None
        self._check(self._L.fb_shard_scan_iter(self._h, int(it), arr, C.c_void_p(int(out_counts))))

    def shardReset(self):
        """After a failed device-side barrier: start flags, error word and pass counter afresh (call on every rank)."""
        self._check(self._L.fb_shard_reset(self._h))

    def shardScanFinish(self, peer_counts, peer_maxprob):
        self._check(self._L.fb_shard_scan_finish(self._h, self._ptr_array(peer_counts), self._ptr_array(peer_maxprob)))

    def shardCounts(self, index, estimate):
        """Summed counts of the gap rows of one unit after shardScanFinish, [estimate][5]."""
        out = np.zeros((int(estimate), 5))
        self._check(self._L.fb_shard_counts(self._h, index, int(estimate), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def ipcExport(self, dev_ptr):
        buf = C.create_string_buffer(64)
        self._check(self._L.fb_ipc_export(self._h, C.c_void_p(int(dev_ptr)), buf))
        return buf.raw

    def ipcOpen(self, handle):
        out = C.c_void_p()
        self._check(self._L.fb_ipc_open(self._h, C.c_char_p(bytes(handle)), C.byref(out)))
        return out.value

    def ipcClose(self, dev_ptr):
        self._check(self._L.fb_ipc_close(self._h, C.c_void_p(int(dev_ptr))))

    def runScan(self):
        self._check(self._L.fb_run_scan(self._h))

    def runFinalize(self):
        self._check(self._L.fb_run_finalize(self._h))

    def scanLiks(self, index):
        a, b = gap_range(self._gaps[index].originalGap)
        out = np.zeros(b - a + 1)
        n = self._check(self._L.fb_scan_liks(self._h, index, out.ctypes.data_as(C.POINTER(C.c_double)), out.size))
        return out[:n]

    def mergeLiks(self, index, liks):
        liks = np.ascontiguousarray(liks, dtype=np.float64)
        self._check(self._L.fb_merge_liks(self._h, index, liks.ctypes.data_as(C.POINTER(C.c_double)), liks.size))

    def hasResult(self, index):
        """True when this context finalized gap `index` (it scanned the winning estimate)."""
        r = _FbGapResult()
        return self._L.fb_result(self._h, index, C.byref(r), None, 0) == FB_OK

    def ownsGap(self, index):
        return bool(self._L.fb_owns_gap(self._h, index))

    def fillGaps(self):
        """Upload + run: host buffers in, filled sequences out (the end-to-end call)."""
        self._check(self._L.fb_run(self._h))

    def upload(self):
        self._check(self._L.fb_upload(self._h))

    def runDevice(self):
        """Everything after the upload (inputs resident in HBM)."""
        self._check(self._L.fb_run_device(self._h))

    def clearGaps(self):
        self._check(self._L.fb_clear_gaps(self._h))
        self._gaps = []

    def stats(self):
        s = FbStats()
        self._check(self._L.fb_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def _result(self, idx):
        r = _FbGapResult()
        cap = 2 * self._gaps[idx].originalGap + 16
        buf = C.create_string_buffer(cap)
        self._check(self._L.fb_result(self._h, idx, C.byref(r), buf, cap))
        return dict(gap_min=r.gap_min, gap_max=r.gap_max, best_estimate=r.best_estimate, best_lik=r.best_lik,
                    total_reads=r.total_reads, discarded=r.discarded, filled_len=r.filled_len,
                    filled=buf.value[:r.filled_len], likelihoods=r.likelihoods)

    def _liks(self, idx):
        a, b = gap_range(self._gaps[idx].originalGap)
        out = np.zeros(b - a + 1)
        n = self._check(self._L.fb_estimate_liks(self._h, idx, out.ctypes.data_as(C.POINTER(C.c_double)), out.size))
        return out[:n]

    def _best_string(self, idx):
        cap = 2 * self._gaps[idx].originalGap + 16
        buf = C.create_string_buffer(cap)
        self._check(self._L.fb_best_string(self._h, idx, buf, cap))
        return buf.value

    def _dump(self, idx, estimate):
        n = (estimate + 2 * self.model.max_distance) * 5
        p, c = np.zeros(n), np.zeros(n)
        dp = C.POINTER(C.c_double)
        self._check(self._L.fb_dump_profile(self._h, idx, int(estimate), p.ctypes.data_as(dp), c.ctypes.data_as(dp)))
        return p.reshape(-1, 5), c.reshape(-1, 5)
