// C++ host-mirror smoke test: the reference's fixtures through host/lpr_solvers.hpp (Program.cs menu paths 1, 2,
// 3, 5 and the cutting plane), checked against the known answers of SURVEY.md Appendix C.  Needs a GPU to RUN;
// tests/test_cpp_host.py compiles and links it everywhere and runs it only under -m gpu.
#include <array>
#include <cstdio>
#include <cstdlib>

#include "../../host/lpr_solvers.hpp"

using namespace LPR_381_Group_V22;
using IO::Constraint;

#define REQUIRE(cond)                                                   \
  do {                                                                  \
    if (!(cond)) {                                                      \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

static void add_cli_bound_rows(int n, std::vector<Constraint>& cons) {  // Program.cs:114-124
  for (int i = 0; i < n; i++) {
    std::vector<double> co(n + 3, 0.0);
    co[i] = 1;
    co[n + 1] = 1;
    cons.push_back({co, "<=", 1.0});
  }
}

int main() {
  // data/TextFile.txt through menu option 1 / 3
  std::vector<double> obj = {2, 3, 3, 5, 2, 4};
  std::vector<Constraint> cons = {{{11, 8, 6, 14, 10, 10}, "<=", 40}};
  add_cli_bound_rows(6, cons);
  Simplex::PrimalSimplexSolver primal(obj, cons);
  primal.Solve();
  REQUIRE(primal.Status == LPR_OPTIMAL);
  REQUIRE(primal.PivotLog.size() == 6 && primal.PivotLog[0] == std::make_pair(5, 3) && primal.PivotLog[5] == std::make_pair(1, 4));
  REQUIRE(primal.FinalZ == 0x1.ecccccccccccdp+3);
  REQUIRE((primal.SolutionVector == std::vector<double>{0, 1, 1, 1, 0.2, 1}));
  REQUIRE((primal.BasicVariables() == std::vector<int>{4, 7, 1, 2, 3, 11, 5}));
  auto bb = IntegerProgramming::BranchAndBoundAdapter::SolveFromPrimal(primal, false, false);
  REQUIRE(bb.second == 15.0 && (bb.first == std::vector<double>{0, 1, 1, 1, 0, 1}));
  // RunBranchAndBound entry (BranchBoundSimplexSolver.cs:1253-1298): same incumbent from the model rows
  {
    IntegerProgramming::BranchBoundSimplexSolver::BranchAndBound rb;
    auto r = rb.RunBranchAndBound(obj, {{11, 8, 6, 14, 10, 10, 40, 0}}, false);
    REQUIRE(r.second == 15.0 && (r.first == std::vector<double>{0, 1, 1, 1, 0, 1}));
    IntegerProgramming::BranchBoundSimplexSolver::DualSimplexSolverBB ds;
    std::vector<std::vector<double>> ge = {{1, 0, 3, 1}, {1, 1, 10, 0}};
    int R = 0, C = 0;
    auto T = ds.FormulateTableau({1, 2}, ge, &R, &C);
    REQUIRE(R == 3 && C == 5 && T[5] == -1.0 && std::signbit(T[6]) && T[9] == -3.0 && T[13] == 1.0);
    REQUIRE(ge[0].size() == 3 && ge[0][0] == -1.0 && ge[1].size() == 3);
  }
  // SensitivityAnalyzer on the device-resident final tableau: add x5 <= 0 (reference sign quirk included)
  {
    SensitivityAnalysis::SensitivityAnalyzer sa(primal.FinalTableau, primal.Rows, primal.Cols, primal.SolutionVector,
                                                primal.FinalZ, primal.BasicVariables());
    REQUIRE((sa.BasicVariables() == std::vector<int>{4, 7, 1, 2, 3, 11, 5}));
    std::vector<double> tech(primal.Cols - 1, 0.0);
    tech[4] = 1.0;
    sa.AddNewConstraintNonInteractive(tech, 0.0);
    REQUIRE(sa.CurrentZ() == 0x1.e2be2be2be2bep+3);
    REQUIRE(sa.SolutionVector()[3] == 0x1.b6db6db6db6dcp-1 && sa.SolutionVector()[4] == 0x1.999999999999ap-2);
    REQUIRE((sa.BasicVariables() == std::vector<int>{4, 7, 1, 2, 3, 11, 5, 10}));
    REQUIRE(sa.ShadowPrices().size() == 8);
  }
  // cutting plane on the final tableau (Appendix C4)
  {
    std::vector<double> o(primal.FinalTableau.begin(), primal.FinalTableau.begin() + primal.Cols);
    std::vector<std::vector<double>> rows;
    for (int i = 1; i < primal.Rows; i++)
      rows.emplace_back(primal.FinalTableau.begin() + (size_t)i * primal.Cols, primal.FinalTableau.begin() + (size_t)(i + 1) * primal.Cols);
    IntegerProgramming::CuttingPlaneSolver cp;
    cp.CuttingPlaneSolution(o, rows);
    REQUIRE(cp.Status == LPR_OPTIMAL && rows.size() == 8 && o.back() == 15.0);
  }
  // README model through menu option 2 (revised; '>=' ignored)
  {
    std::vector<Constraint> c2 = {{{1, 2, 3}, "<=", 10}, {{3, 2, 1}, ">=", 15}};
    Simplex::RevisedPrimalSimplexSolver rev({2, 3, 4}, c2, false);
    rev.Solve();
    REQUIRE(rev.FinalZ == 16.25 && (rev.SolutionVector == std::vector<double>{4.375, 0, 1.875}));
    REQUIRE((rev.BasicVariables() == std::vector<int>{2, 0}));
    // one CaptureSnapshot block per iteration plus the "Optimal" block (RevisedPrimalSimplexSolver.cs:226-246, :124-146)
    REQUIRE(rev.IterationSnapshots.size() == 3);
    REQUIRE(rev.IterationSnapshots[0].rfind("Iteration 1\r\nCurrent Tableau (Revised Simplex)\r\nProblem type: MAX\r\n", 0) == 0);
    REQUIRE(rev.IterationSnapshots[0].find("Entering variable (chosen pre-pivot): x3  (reduced cost pre = 4)") != std::string::npos);
    REQUIRE(rev.IterationSnapshots[2].rfind("Optimal\r\n", 0) == 0);
    REQUIRE(rev.IterationSnapshots[2].find("Basic Variables: x3, x1") != std::string::npos);
    bool threw = false;
    try {
      Simplex::RevisedPrimalSimplexSolver bad({1.0, 0.0}, {{{-1.0, 1.0}, "<=", 1.0}}, false);
      bad.Solve();
    } catch (const std::runtime_error& ex) {
      threw = std::string(ex.what()).find("Unbounded problem") != std::string::npos;
    }
    REQUIRE(threw);
  }
  // knapsack, Program.cs:433-470
  {
    IntegerProgramming::KnapsackBranchBoundSimplex ks(40, {11, 8, 6, 14, 10, 10}, {2, 3, 3, 5, 2, 4});
    const double best = ks.Solve();
    const double dp = IntegerProgramming::KnapsackBranchBoundSolver::Solve(40, {11, 8, 6, 14, 10, 10}, {2, 3, 3, 5, 2, 4});
    REQUIRE(best == 15.0 && std::fabs(dp - best) < 1e-6);
    auto items = ks.GetSelectedItemsOriginal();
    REQUIRE(items.size() == 4 && items[0].Id == 1 && items[3].Id == 5);
  }
  // dual simplex in place
  {
    std::vector<double> o = {1, 2, 3, 0, 0, 0, 0};
    std::vector<std::vector<double>> rows = {{-1, -2, -1, 1, 0, 0, -2}, {-1, -1, -3, 0, 1, 0, -3}, {-2, -1, -1, 0, 0, 1, -4}};
    Simplex::DualSimplexSolver d;
    REQUIRE(d.Solve(o, rows, 10000, true));
    REQUIRE(!Simplex::DualSimplexSolver::AnyNegativeRhs(rows));
  }
  // model input and snapshot text: parse the fixture, build the tableau from the native model,
  // solve, and format the final tableau on the device path and on the host path
  {
    IO::InputFileParser parser;
    parser.ReadInputText("max +2 +3 +3 +5 +2 +4\n+11 +8 +6 +14 +10 +10 <= 40\nbin bin bin bin bin bin");
    REQUIRE(parser.ProblemType == "max" && parser.ObjectiveCoefficients.size() == 6 && parser.Constraints.size() == 1);
    REQUIRE(parser.SignRestrictions.size() == 6 && parser.SignRestrictions[5] == "bin");
    parser.AddCliBoundRows();
    REQUIRE(parser.Constraints.size() == 7 && parser.Constraints[2].Coefficients.size() == 9);
    lpr_tab* t = parser.CreateDeviceTableau();
    int status = 0;
    int64_t np = 0;
    Check(lpr_tab_solve(t, LPR_RULE_PRIMAL, -1, 0, &status, &np, nullptr, 0));
    REQUIRE(status == LPR_OPTIMAL && np == 6);
    double z = 0;
    Check(lpr_tab_objective(t, &z));
    REQUIRE(z == primal.FinalZ);
    const std::string dev = Utilities::TableIterationFormater::Format(t, 6, "Final Tableau (Optimal)");
    const std::string host = Utilities::TableIterationFormater::Format(primal.FinalTableau, primal.Rows, primal.Cols, 6,
                                                                       "Final Tableau (Optimal)");
    REQUIRE(dev == host);
    REQUIRE(dev.find("\nFinal Tableau (Optimal):\r\n") == 0 && dev.find("\tRHS\r\nZ\t") != std::string::npos);
    REQUIRE(dev.find("15.400\t\r\n") != std::string::npos);
    REQUIRE(Utilities::NumFormat::N3(primal.FinalZ) == "15.4" && Utilities::NumFormat::N3(-0.0) == "0");
    lpr_tab_destroy(t);
    IO::InputFileParser missing;
    missing.ReadInputFile("/nonexistent/model.txt");
    REQUIRE(missing.Message.find("can't find your file") != std::string::npos && missing.ObjectiveCoefficients.empty());
    bool threw = false;
    try {
      IO::InputFileParser bad;
      bad.ReadInputText("max 1 2\n1\n+ +");
    } catch (const ArgumentException& ex) {
      threw = std::string(ex.what()).find("IndexOutOfRangeException") != std::string::npos;
    }
    REQUIRE(threw);
  }
  std::printf("cpp host mirror: all checks passed\n");
  return 0;
}
