"""debug: where does cc_graph_end crash? variants in subprocesses"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, faulthandler
faulthandler.enable()
sys.path.insert(0, %r)
import numpy as np
from compute.scala_b200 import cuda
cuda.init(0, streams=int(sys.argv[2]))
T = cuda.Tensor
a, b, c = (T.random([256, 256], seed=s).doCache() for s in (1, 2, 3))
variant = sys.argv[1]
e = T.tanh(a * b + c)
e.doBuffer().release()
cuda.synchronize()
print("begin", variant, flush=True)
with cuda.Graph() as g:
    n = 1 if variant == "one" else 10
    for _ in range(n):
        e.doBuffer().release()
    if variant == "keep":
        out = e.doBuffer()
    if variant == "sum":
        s = (a * b).sum().doBuffer()
print("ended", g.commands, flush=True)
g.launch()
cuda.synchronize()
print("launched", flush=True)
g.release()
print("released", flush=True)
''' % ROOT
for env in ({}, {"CC_PDL": "0"}):
    for variant in ("one", "ten", "keep", "sum"):
        for streams in ("1", "4"):
            r = subprocess.run([sys.executable, "-c", CHILD, variant, streams], env=dict(os.environ, CC_GRAPH_DEBUG="1", **env), capture_output=True, text=True)
            print("==", env, variant, "streams", streams, "rc", r.returncode, "|", r.stdout.replace("\n", " "), "|", r.stderr[-300:].replace("\n", " / "))
