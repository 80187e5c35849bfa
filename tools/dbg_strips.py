import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, krabmaga_b200 as kb
from krabmaga_b200 import strips
from parity_util import NORTH_STAR_DISC, both_params, random_agents
n,w=20000,600.0
agents=random_agents(n,w,w,seed=5)
world=strips.StripWorld(w,w,NORTH_STAR_DISC,10.0,[0,0],n,canonical_order=True,slack=3.0)
for s in world.strips: print(s.rank, s.own_x0,s.own_x1,s.halo_l,s.halo_r,s.dh)
own=strips.owner_of(agents["x"],w,w,NORTH_STAR_DISC,2)
for r,s in enumerate(world.strips):
    m=own==r
    s.upload(agents["id"][m],agents["x"][m],agents["y"][m],agents["ldx"][m],agents["ldy"][m])
    print('uploaded',r,m.sum())
for s in world.strips: s.prepare(); print('prepare enqueued', s.rank)
for s in world.strips:
    try:
        print(s.rank, s.stats())
    except Exception as e: print('ERR',s.rank,e)
_,gp=both_params(exact=0,seed=77)
for i in range(3):
    gp.step=i
    for s in world.strips: s.step_boids(gp)
    for s in world.strips:
        try: print('step',i,s.rank,s.stats())
        except Exception as e: print('ERR',s.rank,e)
