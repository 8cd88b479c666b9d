import os, time, sys, numpy as np
os.environ.setdefault("MOLA_OPTIMIZE_TWIST","false"); os.environ.setdefault("MOLA_INITIAL_VX","8.0")
sys.path.insert(0,"/root/repo")
from mola_lidar_odometry_b200 import synth
from mola_lidar_odometry_b200.api import Context
from mola_lidar_odometry_b200.host_api import LidarOdometry
S=synth.Scene(42); tr=synth.trajectory_T00(130, seed=7)
scans=[S.scan(tr[k],scan_seed=1000+k) for k in range(120)]
for mode in ("default","MLO_PERSISTENT=0"):
    if "=" in mode: os.environ["MLO_PERSISTENT"]="0"
    ctx=Context(0); lo=LidarOdometry(ctx,"/root/repo/pipelines/lidar3d-default.yaml")
    for k in range(20): lo.on_lidar(scans[k],0.1*k)
    ctx.profile_enable(True); ctx.profile_get(True)
    l0=ctx.launch_count; t=time.perf_counter(); its=0
    for k in range(20,120):
        o=lo.on_lidar(scans[k],0.1*k); its+=o.icp_iterations
    dt=time.perf_counter()-t; p=ctx.profile_get(True)
    print(mode,"ms/scan %.3f"%(dt*10),"launches/scan",(ctx.launch_count-l0)/100,"iters/scan",its/100,"filter %.3f icp %.3f map %.3f nn %.3f ms/scan"%(p.filter_1st_ms/100,p.run_icp_ms/100,p.update_local_map_ms/100,p.nn_kernel_ms/100), "nicp", o.n_icp_layer, "nmap", o.n_map_layer)
    ctx.profile_enable(False)
    t=time.perf_counter()
    lo2=LidarOdometry(ctx,"/root/repo/pipelines/lidar3d-default.yaml")
    for k in range(120): lo2.on_lidar(scans[k],0.1*k)
    print(mode,"no-profile ms/scan %.3f"%((time.perf_counter()-t)/120*1e3))
    lo.close(); lo2.close(); ctx.close()
