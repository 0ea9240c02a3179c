"""Loader for the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from runs of
the reference itself)."""
import os

import numpy as np

import oracle_lib as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F0_CASES = ["cfg1_500bp_30x", "lib180_100bp_60x", "multi_gap_err1pct", "tiny_gaps_N_multihit", "plateau_1000bp"]
ALL_CASES = F0_CASES + ["logmode_f1"]


class Case:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.f_mode = int(z["f_mode"])
        ild = np.zeros(int(z["ild_len"]))
        ild[:z["ild_head"].size] = z["ild_head"]
        self.model_args = dict(max_read_len=int(z["max_read_len"]), ild=ild, thr_min=int(z["thr_min"]), thr_max=int(z["thr_max"]),
                               cutoff=int(z["cutoff"]), error_type=z["error_type"], error_pos=z["error_pos"], in_pos=z["in_pos"],
                               del_pos=z["del_pos"], max_distance=int(z["max_distance"]), f_mode=self.f_mode)
        self.D = int(z["max_distance"])
        self.gaps = []
        for gi in range(int(z["n_gaps"])):
            meta = z[f"g{gi}_meta"]
            seqs = bytes(z[f"g{gi}_seqs"]).split(b"\n") if meta.shape[0] else []
            recs = [(int(m[0]), int(m[1]), int(m[2]), s) for m, s in zip(meta, seqs)]
            self.gaps.append(dict(
                window=bytes(z[f"g{gi}_window"]), orig=int(z[f"g{gi}_orig"]), recs=recs, liks=z[f"g{gi}_liks"],
                best_est=int(z[f"g{gi}_best_est"]), best_lik=float(z[f"g{gi}_best_lik"]), best_string=bytes(z[f"g{gi}_best_string"]),
                filled=bytes(z[f"g{gi}_filled"]), counts=tuple(int(v) for v in z[f"g{gi}_counts"]), rows=z[f"g{gi}_rows"]))

    def oracle_model(self):
        return O.Model(**self.model_args)

    def figbird_model(self):
        from gapprober_b200 import FigbirdModel
        return FigbirdModel(**self.model_args)
