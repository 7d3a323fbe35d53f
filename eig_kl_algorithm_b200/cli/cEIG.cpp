// cEIG -- drop-in for the reference's cEIG executable (cEIG.cpp:138-237): same argv, same output
// file pre_saved_EIG/<base>_out.txt, same error convention ("Error: <what>" on stderr, exit code 1).
// The work runs on the GPU through libeigkl (assembly -> Lanczos -> median/sides).
#include <chrono>
#include <iostream>
#include "cli_common.h"

int main(int argc, char *argv[]) {
  auto t0 = std::chrono::high_resolution_clock::now();
  StageClock clk;
  eigkl_handle *h = nullptr;
  auto fail = [&](const std::string &what) {
    std::cerr << "Error: " << what << std::endl;        // cEIG.cpp:231-234
    if (h) eigkl_destroy(h);
    return 1;
  };
  if (argc != 2) return fail("Usage: ./EIG <input_file>");           // cEIG.cpp:143-145
  create_dir("results");                                              // cEIG.cpp:148-149
  create_dir("pre_saved_EIG");
  const std::string filename = argv[1];
  const std::string outfile = "pre_saved_EIG/" + base_name(filename) + "_out.txt";   // cEIG.cpp:162-164
  eigkl_opts o{};
  o.struct_size = sizeof(o);
  o.device = device_from_env();
  if (eigkl_create(&h, &o) != EIGKL_OK) return fail(eigkl_last_error(nullptr));
  clk.tick("eigkl_create (CUDA context)");
  std::cout << "\n============= Initialization =============\n";
  std::cout << "Backend: libeigkl (CUDA, sm_100a), ABI " << eigkl_abi_version() << std::endl;
  if (eigkl_load_hgr(h, filename.c_str()) != EIGKL_OK) return fail(eigkl_last_error(h));
  int32_t nodes = 0, nets = 0;
  eigkl_get_sizes(h, &nodes, &nets, nullptr);
  std::cout << "\nProblem Size:\n  - Nets: " << nets << "\n  - Nodes: " << nodes << "\n";
  clk.tick("eigkl_load_hgr");
  std::cout << "\nInitializing sparse matrix...\n";
  if (eigkl_assemble_laplacian(h) != EIGKL_OK) return fail(eigkl_last_error(h));
  clk.tick("eigkl_assemble_laplacian");
  std::cout << "Computing eigenvalues...\n";
  double lambda2 = 0;
  if (eigkl_fiedler(h, &lambda2, nullptr) != EIGKL_OK) return fail(std::string("Eigenvalue computation failed: ") + eigkl_last_error(h));
  clk.tick("eigkl_fiedler");
  std::cout << "\nWriting results...\n";
  if (eigkl_write_eig(h, outfile.c_str()) != EIGKL_OK) return fail(eigkl_last_error(h));
  clk.tick("eigkl_write_eig");
  eigkl_stats st{};
  st.struct_size = sizeof(st);
  eigkl_get_stats(h, &st);
  auto t1 = std::chrono::high_resolution_clock::now();
  std::cout << "\n============= Summary =============\n";
  std::cout << "lambda2: " << lambda2 << "  (" << st.matvecs << " matvecs, " << st.restarts << " restarts, ncv " << st.ncv << ")\n";
  std::cout << "GPU time: assembly " << st.ms_assemble_laplacian << " ms, solve " << st.ms_fiedler << " ms\n";
  std::cout << "Execution time: " << std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count() / 1000.0 << " seconds\n";
  std::cout << "Results written to: " << outfile << "\n";
  std::cout << "================================\n\n";
  eigkl_destroy(h);
  return 0;
}
