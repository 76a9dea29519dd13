import sys; sys.path.insert(0,'/root/repo')
import numpy as np, json
import recastnavigation_b200 as rcb
from recastnavigation_b200 import geom
ag=geom.Agent(cs=0.2,ch=0.2)
v,t=geom.terrain_with_boxes(side_m=2828.4*0.2**0.5,n_boxes=200000)
a=geom.mark_walkable_triangles(v,t,ag.slope)
bmin,bmax=geom.calc_bounds(v)
f,tw,th=geom.tile_fields(bmin,bmax,ag.cs,ag.ch,64,ag.walkable_radius+3)
ts,tl=geom.tile_triangle_lists(v,t,f,tw,th)
b=rcb.Batch(0); b.set_geometry(v,t,a); b.set_fields(f,ts,tl)
c=b.run(ag.walkable_height,ag.walkable_climb,ag.walkable_radius,ag.walkable_climb)
r=b.field_results()
cc,_=b.get_solid()
print('fields',f.size,'pairs',c.pairs,'literalPairs',c.literalPairs,'slots',c.fragmentSlots,'frags',c.fragments,'S',c.solidSpans,'K',c.compactSpans)
print('S per field: mean %.0f max %d p99 %.0f'%(r['solidSpans'].mean(), r['solidSpans'].max(), np.percentile(r['solidSpans'],99)))
print('K per field: mean %.0f max %d p99 %.0f'%(r['compactCount'].mean(), r['compactCount'].max(), np.percentile(r['compactCount'],99)))
print('tris per field: mean %.0f max %d'%(np.diff(ts.astype(np.int64)).mean(), np.diff(ts.astype(np.int64)).max()))
print('max col count',cc.max(), 'maxdist', r['maxDistance'].max(), r['maxDistance'].mean())
