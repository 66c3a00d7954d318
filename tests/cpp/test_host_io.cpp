// CPU-only driver of the host-code part of the C++ mirror (host/lpr_solvers.hpp): IO::InputFileParser,
// AddUpperBoundConstraints, IO::OutputFileWrite, Utilities::TableIterationFormater / NumFormat.  tests/test_cpp_host.py
// feeds it the inputs of tests/golden/reference_run.json (what the reference's own classes did with them, executed by
// oracle/csharp) and compares what it prints and writes.  No GPU is needed: these entry points are host code.
//
//   test_host_io <model.txt> <add_upper_bound_rows 0|1> <out.txt> <solver> <timestamp> <z hex> <n_x> <x hex>... <n_snap> <snapfile>...
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>
#include <vector>

#include "../../host/lpr_solvers.hpp"

using namespace LPR_381_Group_V22;

static std::string slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  return std::string(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
  if (argc < 8) return 2;
  int a = 1;
  const std::string text = slurp(argv[a++]);
  const bool add_bounds = std::atoi(argv[a++]) != 0;
  const std::string out = argv[a++], solver = argv[a++], stamp = argv[a++];
  const double z = std::strtod(argv[a++], nullptr);
  std::vector<double> x(std::atoi(argv[a++]));
  for (auto& v : x) v = std::strtod(argv[a++], nullptr);
  std::vector<std::string> snaps(std::atoi(argv[a++]));
  for (auto& s : snaps) s = slurp(argv[a++]);

  IO::InputFileParser p;
  try {
    p.ReadInputText(text);
  } catch (const std::exception& e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 0;
  }
  std::printf("MESSAGE %s\n", p.Message.c_str());
  if (p.ProblemType.empty()) return 0;
  const size_t before = p.Constraints.size();
  if (add_bounds) p.AddUpperBoundConstraints();
  std::printf("TYPE %s\nOBJ", p.ProblemType.c_str());
  for (double c : p.ObjectiveCoefficients) std::printf(" %a", c);
  std::printf("\n");
  for (size_t i = 0; i < p.Constraints.size(); i++) {
    std::printf(i < before ? "ROW" : "ADDED");
    for (double c : p.Constraints[i].Coefficients) std::printf(" %a", c);
    std::printf(" | %s | %a\n", p.Constraints[i].Relation.c_str(), p.Constraints[i].RHS);
  }
  std::printf("SIGNS");
  for (auto& s : p.SignRestrictions) std::printf(" [%s]", s.c_str());
  std::printf("\n");
  IO::OutputFileWrite::WriteFullResults(out, solver, p, snaps, z, x.empty() ? nullptr : &x, false, stamp);
  IO::OutputFileWrite::WriteSnapshotsOnly(out, solver + " (again)", snaps, z, x.empty() ? nullptr : &x, true, stamp);
  std::vector<double> tab = {0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 0.0005, -0.0005, 0.0015, 2.0005, 1234.5675, 1e-13};
  std::printf("TABLE_BEGIN\n%sTABLE_END\n", Utilities::TableIterationFormater::Format(tab, 3, 4, 2, "T").c_str());
  std::printf("N3 %s %s %s\n", Utilities::NumFormat::N3(2.0005).c_str(), Utilities::NumFormat::N3(-1e-13).c_str(),
              Utilities::NumFormat::N3(1e15).c_str());
  return 0;
}
