import os, sys, ctypes
sys.path.insert(0, os.getcwd())
import bench
from gapprober_b200 import FigbirdContext, FigbirdModel, read_gap_info, read_scaffolds, figbird
wd = bench.prepare_workload('cfg2_rhodobacter', 2)
model = FigbirdModel.load(os.path.join(wd, 'fb_model.bin'))
with FigbirdContext(model) as fb:
    fb.add_workdir(wd)
    fb.upload()
    fb.runDevice(); fb.runDevice()
    L = figbird.load_library()
    L.fb_prof_dump()
    fb.runDevice()
    print('em ms', fb.stats()['em_kernel_ms'])
    L.fb_prof_dump()
