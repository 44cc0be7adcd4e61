import sys
sys.path.insert(0,'/root/repo')
from solver2d_b200 import capi, device, scenes
P = capi.Solver2D(device.LIB_PATH); dev = device.Device()
sc = scenes.pyramid(P, "TGS_Soft", base_count=447)
dw = device.DeviceWorld.attach(dev, sc.world)
out=[]
for i in range(60):
    sc.step(1/60, 4, 2, True)
    out.append(dw.counters().movedCount)
print(out)
