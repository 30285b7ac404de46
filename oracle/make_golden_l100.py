"""oracle/make_golden_l100.py -- regenerates tests/golden/tc_l100_r20.json (build container only; ~90 s on 8 cores).
Truth = exhaustive switching (oracle/pt_oracle.c) on BASELINE-config-2 sized instances with two guest labels swapped.
Instances checked: the first 40 plus every one the kernel-logic simulator (tests/host_sim) calls displayed."""
import sys, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,'/root/repo')
import phylo_tools_b200 as pt
from oracle import pt_oracle as O
from concurrent.futures import ProcessPoolExecutor
count, seed, kind = 400, 45568, 1
p = pt.Packer(); p.add_generated(count, 100, 20, seed, kind)
pairs=[p.newick(i) for i in range(count)]
# the instances the simulator says are displayed (cheap to confirm: stop at first) + 40 others (full 2^20 enumeration)
import subprocess
got = subprocess.run([os.path.join(ROOT,'tests','host_sim','build','sim_tc'),'gen',str(count),'100','20',str(seed),str(kind),'0'],capture_output=True,text=True).stdout.strip()
idx = sorted(set([i for i,c in enumerate(got) if c=='1'] + list(range(40))))
def f(i): return O.tc_bruteforce_newick(pairs[i][0], pairs[i][1], stop_at_first=True, max_switchings=1<<24)[0]
if __name__=='__main__':
    with ProcessPoolExecutor(8) as ex: truth=list(ex.map(f, idx))
    json.dump(dict(source="oracle/pt_oracle.c exhaustive switching over pt_packer_add_generated(count, 100, 20, seed, kind)", count=count, seed=seed, kind=kind, index=idx, truth=truth), open('/root/repo/tests/golden/tc_l100_r20.json','w'))
    print(len(idx), sum(truth), [got[i] for i in idx]==[str(t) for t in truth])
