"""FTL streaming microbench on its own (for ncu): python scratch/ftl_diag.py"""
import json, sys
import torch
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.ftl_stream(torch.device("cuda", 0)), indent=1))
