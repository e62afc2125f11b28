import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, configs
import optika_b200 as optika
from optika_b200 import named as na
from oracle import raytrace as ora
system = configs.newtonian(num_field=10, num_pupil=32)
_, rays = system._input(None,None,None,None,False,False)
r0,_ = configs.flatten_rays(rays); r0={k:v.reshape(-1) for k,v in r0.items()}
surfaces = system.surfaces_all
want = ora.accumulate_rays(surfaces, r0)
dev = optika.propagators.accumulate_rays(surfaces, rays, axis="surface")
d, shape_ = configs.flatten_rays(dev)
k = list(shape_).index("surface")
got = {n: np.moveaxis(v,k,0).reshape(shape_["surface"],-1) for n,v in d.items()}
for s in range(6):
    for n in ('px','py','pz','dx','dy','dz'):
        e = np.abs(got[n][s]-want[n][s])
        i = np.nanargmax(e)
        print(s, n, 'max abs err %.3e'%e[i], 'at', i, 'val', want[n][s][i])
# single-surface: feed oracle state after surface 2 into GPU primary only
st = {n: want[n][2] for n in want}
ax='ray'
rin = optika.rays.RayVectorArray(
    wavelength=na.ScalarArray(st['wavelength'],ax),
    position=na.Cartesian3dVectorArray(*[na.ScalarArray(st[c],ax) for c in ('px','py','pz')]),
    direction=na.Cartesian3dVectorArray(*[na.ScalarArray(st[c],ax) for c in ('dx','dy','dz')]),
)
o = surfaces[3].propagate_rays(rin)
w1 = ora.surface_propagate(surfaces[3], st)
for n,c in (('px',o.position.x),('py',o.position.y),('pz',o.position.z),('dx',o.direction.x)):
    e = np.abs(c.ndarray - w1[n]); i=np.nanargmax(e)
    print('primary alone', n, 'max abs err %.3e'%e[i], 'bitwise equal frac', (c.ndarray==w1[n]).mean())
# intercept only
from optika_b200 import _engine
oi = _engine.sag_intercept(surfaces[3].sag, rin, attenuate=False)
loc = ora._rays_transform(surfaces[3].transformation, st, inverse=True)
wi = ora.sag_intercept(surfaces[3].sag, st)
print('NOTE sag-only uses untransformed rays')
loc_rays = optika.rays.RayVectorArray(
    wavelength=na.ScalarArray(loc['wavelength'],ax),
    position=na.Cartesian3dVectorArray(*[na.ScalarArray(loc[c],ax) for c in ('px','py','pz')]),
    direction=na.Cartesian3dVectorArray(*[na.ScalarArray(loc[c],ax) for c in ('dx','dy','dz')]),
)
oi = _engine.sag_intercept(surfaces[3].sag, loc_rays, attenuate=False)
wi = ora.sag_intercept(surfaces[3].sag, loc)
for n,c in (('px',oi.position.x),('py',oi.position.y),('pz',oi.position.z)):
    e = np.abs(c.ndarray - wi[n]); i=np.nanargmax(e)
    print('intercept alone', n, 'max abs err %.3e'%e[i], 'bitwise equal frac', (c.ndarray==wi[n]).mean())
