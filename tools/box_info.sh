#!/bin/bash
# What the GPU box looks like from the host side (development aid for the multi-GPU host-to-host path).
echo "== nproc / lscpu"; nproc; lscpu | head -40
echo "== numa"; ls /sys/devices/system/node/ | tr '\n' ' '; echo; for n in /sys/devices/system/node/node*; do echo "$n cpus=$(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
echo "== meminfo"; grep -E "MemTotal|MemFree|HugePages|Hugepagesize|AnonHuge" /proc/meminfo
echo "== thp"; cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag 2>&1
echo "== cgroup cpu"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; taskset -p $$
echo "== ulimit -l"; ulimit -l
echo "== gpus"; nvidia-smi -L
echo "== topo"; nvidia-smi topo -m 2>&1
echo "== pci numa"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d numa=$(cat $d/numa_node) class=$(cat $d/class) link=$(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; fi; done
echo "== lspci tree"; lspci -tv 2>/dev/null | head -80
echo "== nvidia-smi pcie"; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
echo "== virt"; systemd-detect-virt 2>/dev/null; grep -m1 hypervisor /proc/cpuinfo | head -1; dmesg 2>/dev/null | grep -i -E "iommu|DMAR" | head -5
