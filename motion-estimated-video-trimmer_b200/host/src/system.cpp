// system.cpp — see system.hpp.
#include "motion_trim/system.hpp"

#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <cctype>
#include <fstream>

namespace motion_trim {

std::vector<int> get_available_cpus() {
  std::vector<int> out;
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof set, &set) == 0)
    for (int c = 0; c < CPU_SETSIZE; ++c)
      if (CPU_ISSET(c, &set)) out.push_back(c);
  if (out.empty()) out.push_back(0);
  return out;
}

bool pin_thread_to_cpus(const std::vector<int>& cpus) {
  if (cpus.empty()) return false;
  cpu_set_t set;
  CPU_ZERO(&set);
  for (int c : cpus)
    if (c >= 0 && c < CPU_SETSIZE) CPU_SET(c, &set);
  return pthread_setaffinity_np(pthread_self(), sizeof set, &set) == 0;
}

std::vector<int> pci_local_cpus(const std::string& pci_bus_id) {
  std::vector<int> out;
  std::string id = pci_bus_id;
  for (char& ch : id) ch = (char)std::tolower((unsigned char)ch);
  std::ifstream f("/sys/bus/pci/devices/" + id + "/local_cpulist");
  std::string s;
  if (!f || !std::getline(f, s)) return out;
  size_t i = 0;  // "0-15,32-47"
  while (i < s.size()) {
    size_t j = i;
    while (j < s.size() && std::isdigit((unsigned char)s[j])) ++j;
    if (j == i) break;
    const int a = std::stoi(s.substr(i, j - i));
    int b = a;
    if (j < s.size() && s[j] == '-') {
      size_t k = ++j;
      while (k < s.size() && std::isdigit((unsigned char)s[k])) ++k;
      if (k == j) break;
      b = std::stoi(s.substr(j, k - j));
      j = k;
    }
    for (int c = a; c <= b; ++c) out.push_back(c);
    i = j + 1;
  }
  return out;
}

std::string cpu_list_string(const std::vector<int>& cpus) {
  std::string s;
  for (size_t i = 0; i < cpus.size(); ++i) s += (i ? "," : "") + std::to_string(cpus[i]);
  return s;
}

}  // namespace motion_trim
