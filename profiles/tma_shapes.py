"""Correctness sweep (against float64) of the grouped Linear forward / data-gradient products over the layer shapes of the
generators and discriminators, small group counts included:   python profiles/tma_shapes.py"""
import os
import sys

sys.argv = [sys.argv[0]]
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pair_check.py")).read()
exec(src.split('print("CGL_TUNE =')[0])
worst = 0.0
for G in (1, 3):
    for rows in (100, 300, 37):
        for din, dout in [(100, 128), (128, 256), (256, 512), (512, 1024), (1024, 784), (784, 512), (512, 256), (100, 256), (256, 128)]:
            worst = max(worst, check(G, rows, din, dout))
            torch.cuda.synchronize()
print("worst error", worst)
