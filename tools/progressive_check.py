import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from megaverse_b200 import capi
import helpers
# progressive download must deliver the same frames as zero-copy
a=capi.Engine("ObstaclesHard",64,2,128,72,num_threads=4,depth=True); b=capi.Engine("ObstaclesHard",64,2,128,72,num_threads=4,depth=True)
a.seed(5); b.seed(5); b.set_option("zero_copy",0); b.set_option("progressive",1)
a.reset(); b.reset()
rng=np.random.default_rng(0)
for t in range(120):
    acts=helpers.purposeful_actions(rng,128,t); a.step(acts); b.step(acts)
    assert np.array_equal(np.array(a.obs()),np.array(b.obs())), t
    assert np.array_equal(np.array(a.depth()),np.array(b.depth())), t
print("progressive == zero-copy over 120 steps; faults", a.faults(), b.faults())
