// system.hpp — CPU discovery and thread pinning of the host mirror (reference include/motion_trim/system.hpp,
// src/system.cpp:166-225: get_available_cpus, pin_thread_to_cpus). The cgroup quota detection of the reference is
// out of scope; the CPUs are the ones the process may run on (sched_getaffinity honours cpusets and taskset).
#pragma once

#include <string>
#include <vector>

namespace motion_trim {

std::vector<int> get_available_cpus();
// Pins the calling thread to `cpus` (pthread_setaffinity_np, like src/system.cpp:211-225). Empty set: no-op, false.
bool pin_thread_to_cpus(const std::vector<int>& cpus);
// CPUs local to the PCI device `pci_bus_id` ("0000:3b:00.0", from mscan_device_pci_bus_id) per sysfs; empty if unknown.
std::vector<int> pci_local_cpus(const std::string& pci_bus_id);
std::string cpu_list_string(const std::vector<int>& cpus);

}  // namespace motion_trim
