"""Regression tool (no GPU): SHA-1 of the generated kernel text, plan kind and launch count of 250 random graphs, the five BASELINE
configurations and two dense windows, for the checkout given as argv[1], written to argv[2].  Run it on two checkouts (e.g. `git worktree
add /tmp/wt <commit>` + build there) and diff the JSON files: identical output = the device-side code the generator emits did not
change.  Used at the end of round 1 to show that the host-side work after the last full GPU test run (launch path, SmallVec, JIT outside
the lock, opt-in lowerings) left every default plan byte-identical to the tested commit."""
import sys, hashlib, json, os
root=sys.argv[1]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root,'tests'))
import numpy as np
from compute.scala_b200 import cuda
assert root in cuda._L()._name, cuda._L()._name
from test_fuzz_compile_only import LazyGen, DIMS_WIDE
out={}
def rec(name, k):
    src=k.source
    out[name]=[hashlib.sha1(src.encode()).hexdigest(), k.info.kind, k.info.n_launches]
for seed in list(range(90000, 90150))+list(range(95000,95100)):
    gen = LazyGen(cuda, seed, dims=DIMS_WIDE if seed % 2 else (1, 2, 3, 4, 5, 8, 12), max_rank=2 if seed % 2 else 4)
    try:
        p = gen.expr(depth=2)
        if int(np.prod(p.shape)) > 2_000_000: continue
        rec(f"fuzz{seed}", p.g.compile())
    except cuda.ComputeCudaError as e:
        out[f"fuzz{seed}"]=["ERR "+str(e)[:60]]
T=cuda.Tensor
def chain(parts, f=lambda a,b:a+b):
    acc=parts[0]
    for q in parts[1:]: acc=f(acc,q)
    return acc
a,b,c=(T.random([1024,1024],seed=s) for s in (1,2,3))
rec("C1", T.tanh(a*b+c).compile())
A,B,Cc=(T.random([16384,16384],seed=s) for s in (1,2,3))
t_=A*B+Cc; u=T.exp(t_); v=T.log(u+A); w=T.tanh(v*B); rec("C2",(w+Cc).compile())
x=T.random([16384,16384],seed=5)
rec("C3full", x.sum().compile()); rec("C3ax0", chain(x.split(0)).compile()); rec("C3ax1", chain(x.split(1)).compile())
t3=T.random([512,512,512],seed=7); m2=T.random([512,512],seed=8)
rec("C4a", t3.permute([2,0,1]).translate([3,-5,7]).compile()); rec("C4b", m2.broadcast([512,512,512]).compile())
rec("C4c", m2.reshape([1,512,512]).broadcast([512,512,512]).compile()); rec("C4d", T.join(t3.split(1)).compile())
n=8192
Am,Bm=T.random([n,n],seed=9),T.random([n,n],seed=10)
# matmul2 at smaller size for speed of graph building
n=512
Am,Bm=T.random([n,n],seed=9),T.random([n,n],seed=10)
rec("C5_512", chain((Am.broadcast([n,n,n])*Bm.reshape([1,n,n]).broadcast([n,n,n])).split(1)).compile())
def box(nn, r):
    xx=T.random([nn,nn],seed=1)
    return chain([xx.translate([dy,dx]) for dy in range(-r,r+1) for dx in range(-r,r+1)])
rec("box3", box(4096,1).compile()); rec("box5", box(4096,2).compile())
json.dump(out, open(sys.argv[2],'w'), indent=0)
print(len(out), "plans")
