// kl_main.h -- body shared by the cKL and gKL drop-ins (cKL.cpp:424-468, gKL.cu:672-713).
#pragma once
#include <algorithm>
#include <chrono>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <random>
#include <vector>
#include "cli_common.h"

// gkl_flavour: usage text on stderr when argc < 2 (gKL.cu:673-676); cKL prints it on stdout when
// argc is not 2 or 3 (cKL.cpp:431-434).  Both return 1.
// Extensions (none of them changes what the reference's own argv forms do):
//   --seed <n>             seed of the random initial partition (cKL.cpp:175-193 shuffles with an unseeded
//                          mt19937; EIGKL_SEED in the environment does the same)
//   --rollback             keep only the swaps up to the best cut (the reference tracks it, cKL.cpp:363, and stops there)
//   --partition-out <file> write the final partition, "<node>\t<side>" per line (default name with --rollback alone:
//                          results/<base>_KL_partition[_EIG].txt)
inline int kl_main(int argc_all, char *argv_all[], bool gkl_flavour) {
  create_dir("results");                                              // cKL.cpp:428-429
  create_dir("pre_saved_EIG");
  bool rollback = false, have_seed = false;
  unsigned long seed_value = 0;
  std::string partition_out;
  std::vector<char *> pos;
  for (int i = 0; i < argc_all; ++i) {
    const std::string a = argv_all[i];
    if (i > 0 && a == "--rollback") rollback = true;
    else if (i > 0 && a == "--seed" && i + 1 < argc_all) { have_seed = true; seed_value = strtoul(argv_all[++i], nullptr, 10); }
    else if (i > 0 && a == "--partition-out" && i + 1 < argc_all) partition_out = argv_all[++i];
    else pos.push_back(argv_all[i]);
  }
  const int argc = (int)pos.size();
  char **argv = pos.data();
  if (gkl_flavour ? (argc < 2) : (argc != 2 && argc != 3)) {
    (gkl_flavour ? std::cerr : std::cout) << "Usage: " << argv[0] << " <input_file> [-EIG]" << std::endl;
    return 1;
  }
  const std::string input_file = argv[1];
  const std::string base = base_name(input_file);
  std::string fout_name = "results/" + base + "_KL_CutSize_output.txt";        // cKL.cpp:438
  std::string eig_file;
  bool eig_init = false;
  if (argc >= 3 && strcmp(argv[2], "-EIG") == 0) {                              // cKL.cpp:440-444
    eig_init = true;
    eig_file = "pre_saved_EIG/" + base + "_out.txt";
    fout_name = "results/" + base + "_KL_CutSize_EIG_output.txt";
  }
  StageClock clk;
  eigkl_handle *h = nullptr;
  auto fail = [&](const std::string &what) {
    std::cerr << what << std::endl;
    if (h) eigkl_destroy(h);
    return 1;
  };
  eigkl_opts o{};
  o.struct_size = sizeof(o);
  o.device = device_from_env();
  if (eigkl_create(&h, &o) != EIGKL_OK) return fail(std::string("Error occurred: ") + eigkl_last_error(nullptr));
  clk.tick("eigkl_create (CUDA context)");
  std::cout << "\n============= Reading Input File ==============\n";
  if (eigkl_load_hgr(h, input_file.c_str()) != EIGKL_OK) {
    const std::string e = eigkl_last_error(h);
    return fail(e.rfind("Error opening", 0) == 0 ? "Error opening file: " + input_file : "Error occurred: " + e);   // cKL.cpp:87-90
  }
  clk.tick("eigkl_load_hgr");
  int32_t nodes = 0, nets = 0;
  eigkl_get_sizes(h, &nodes, &nets, nullptr);
  std::cout << "Circuit Statistics\n  - Total Nets : " << nets << "\n  - Total Nodes: " << nodes << "\n";
  if (eigkl_assemble_kl_graph(h) != EIGKL_OK) return fail(std::string("Error occurred: ") + eigkl_last_error(h));
  clk.tick("eigkl_assemble_kl_graph");
  std::cout << "\n\n=========== Starting KL Algorithm =============\n";
  int64_t n0 = 0, n1 = 0;
  if (eig_init) {
    if (eigkl_load_eig(h, eig_file.c_str()) != EIGKL_OK) return fail(eigkl_last_error(h));   // "Error: EIG file not found", cKL.cpp:157-160
    std::vector<uint8_t> side((size_t)nodes);
    eigkl_get_partition(h, side.data());
    n1 = std::accumulate(side.begin(), side.end(), (int64_t)0);
    n0 = nodes - n1;
  } else {
    // random half split, cKL.cpp:175-193 (unseeded mt19937 shuffle in the reference; EIGKL_SEED makes it repeatable)
    std::vector<int32_t> ids((size_t)nodes);
    std::iota(ids.begin(), ids.end(), 0);
    const char *seed_env = getenv("EIGKL_SEED");
    const unsigned seed = have_seed ? (unsigned)seed_value : seed_env ? (unsigned)strtoul(seed_env, nullptr, 10) : std::random_device{}();
    std::mt19937 gen(seed);
    std::shuffle(ids.begin(), ids.end(), gen);
    n0 = nodes / 2; n1 = nodes - n0;
    if (eigkl_set_partition_ordered(h, ids.data(), n0, ids.data() + n0, n1) != EIGKL_OK) return fail(std::string("Error occurred: ") + eigkl_last_error(h));
  }
  clk.tick("initial partition");
  std::cout << "Partition sizes - Left: " << n0 << " Right: " << n1 << std::endl;
  const int64_t cap = std::min(n0, n1) + 1;
  std::vector<float> cut((size_t)cap), gain((size_t)cap);
  std::vector<int32_t> a((size_t)cap), b((size_t)cap);
  eigkl_trace tr{};
  tr.capacity = cap; tr.cut = cut.data(); tr.gain = gain.data(); tr.node1 = a.data(); tr.node2 = b.data();
  auto t0 = std::chrono::high_resolution_clock::now();
  if (eigkl_kl_run(h, &tr) != EIGKL_OK) return fail(std::string("Error occurred: ") + eigkl_last_error(h));
  auto t1 = std::chrono::high_resolution_clock::now();
  clk.tick("eigkl_kl_run");
  if (eigkl_write_trace(fout_name.c_str(), &tr) != EIGKL_OK) return fail("Error: Cannot open output file");   // cKL.cpp:296-299
  clk.tick("eigkl_write_trace");
  eigkl_stats st{};
  st.struct_size = sizeof(st);
  eigkl_get_stats(h, &st);
  float best = cut[0];
  for (int64_t i = 1; i <= tr.swaps; ++i) best = std::min(best, cut[(size_t)i]);
  if (rollback || !partition_out.empty()) {
    if (partition_out.empty()) partition_out = "results/" + base + (eig_init ? "_KL_partition_EIG.txt" : "_KL_partition.txt");
    int64_t kept = tr.swaps;
    float kept_cut = cut[(size_t)tr.swaps];
    if (rollback && eigkl_kl_rollback(h, &kept, &kept_cut) != EIGKL_OK) return fail(std::string("Error occurred: ") + eigkl_last_error(h));
    if (eigkl_write_partition(h, partition_out.c_str()) != EIGKL_OK) return fail("Error: Cannot open output file");
    std::cout << "\nPartition written to " << partition_out << " (after swap " << kept << ", cut " << kept_cut << ")\n";
  }
  std::cout << "\nInitial Partition Information:\n  - Left partition size: " << n0 << "\n  - Right partition size: " << n1
            << "\n  - Initial cut size: " << cut[0] << "\n";
  std::cout << "\n\n=============== Final Results =================\n";
  std::cout << std::left << std::setw(24) << "Total iterations" << ": " << tr.swaps << "\n";
  std::cout << std::left << std::setw(24) << "Initial cut size" << ": " << std::fixed << std::setprecision(2) << cut[0] << "\n";
  std::cout << std::left << std::setw(24) << "Best cut size achieved" << ": " << best << "\n";
  std::cout << std::left << std::setw(24) << "Overall improvement" << ": " << 100.0f * (1.0f - best / cut[0]) << "%\n";
  std::cout << std::left << std::setw(24) << "Total runtime" << ": "
            << std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count() / 1000.0 << " seconds"
            << " (GPU: setup " << st.ms_kl_setup << " ms, swap loop " << st.ms_kl_loop << " ms)\n";
  std::cout << std::left << std::setw(24) << "Trace written to" << ": " << fout_name << "\n";
  eigkl_destroy(h);
  clk.tick("eigkl_destroy");
  return 0;
}
