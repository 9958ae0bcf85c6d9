nvidia-smi topo -m 2>&1 | head -30
echo "--- nodes"; ls /sys/devices/system/node/ 2>&1 | head; cat /sys/devices/system/node/online 2>&1
for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist 2>/dev/null) $(grep MemTotal $n/meminfo 2>/dev/null); done
echo "--- affinity"; python -c "import os;print(len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:40], os.cpu_count())"
echo "--- gpu numa"; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302\|^0x0300" $d/class 2>/dev/null && grep -q 0x10de $d/vendor; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done
echo "--- cgroup"; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null; cat /sys/fs/cgroup/cpu.max 2>/dev/null
grep -i "Mems_allowed_list\|Cpus_allowed_list" /proc/self/status
which numactl; lscpu | head -25
