import sys, time
import os; HERE=os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np, torch
import junctiontree as jt, jt_workloads as wl
for net in (wl.dag37(), wl.ising(16), wl.large_state_tree(), wl.dag500()):
    t0=time.perf_counter()
    tree = jt.create_junction_tree(net['factors'], net['sizes'], order=net.get('order'))
    t1=time.perf_counter()
    out = tree.propagate(net['values'])
    t2=time.perf_counter()
    n=5
    for _ in range(n): out = tree.propagate(net['values'])
    t3=time.perf_counter()
    print(net['name'], 'compile %.3f s first call %.3f s steady %.3f ms'%(t1-t0, t2-t1, (t3-t2)/n*1e3), 'Z', float(out[0].sum()))
