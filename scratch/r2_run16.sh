#!/bin/bash
# round 2, GPU call 16 (8 GPUs): topology, per-GPU PCIe RX during the end-to-end leg, the N=8 bench line
cd "$(dirname "$0")/.."
O=gpurun_out
{
nvidia-smi topo -m
echo "---- lscpu"; lscpu | grep -i -E "model name|socket|numa|^cpu\(s\)|thread|core"
echo "---- numa_node of each GPU"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d $(cat $d/numa_node) $(cat $d/local_cpulist)"; fi; done
echo "---- free"; free -g | head -2
} > $O/r2p_topology_8gpu.txt 2>&1
head -30 $O/r2p_topology_8gpu.txt
(nvidia-smi dmon -s t -d 1 -c 150 > $O/r2p_dmon_pcie_8gpu.txt 2>&1 &)
(time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus 8 --steps 20 --warmup 3 --sub-scans 300) > $O/r2p_bench_n8.json 2> $O/r2p_bench_n8.err
tail -5 $O/r2p_bench_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/r2p_bench_n8.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('N=8 value',round(d['value']),'e2e',d['e2e'],'subs',{k:round(v['value']) for k,v in d.get('sub_records',{}).items()})
PY
tail -20 $O/r2p_dmon_pcie_8gpu.txt
