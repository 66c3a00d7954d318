// host/lpr_solvers.hpp -- C++ host-side mirror of the reference's solver classes over the C ABI.
//
// The reference is compiled C# (no .NET toolchain in this image), so the host layer above liblprb200 is
// also provided in C++: same class names, member names, argument meaning and error behaviour as
//   Simplex/PrimalSimplexSolver.cs, Simplex/RevisedPrimalSimplexSolver.cs, Simplex/PrimalSimplexSolver2.cs,
//   Simplex/DualSimplex.cs, IntegerProgramming/CuttingPlaneSolver.cs, IntegerProgramming/BranchAndBoundAdapter.cs,
//   BranchBoundSimplexSolver.{DualSimplexSolverBB, BranchAndBound} (RunBranchAndBound entry),
//   SensitivityAnalysis/SensitivityAnalyzer.cs (re-optimisation members) and the knapsack classes of
//   Program.cs:430-471.
// Header only; link with -llprb200.  No arithmetic happens here: every pivot runs on the GPU.
#pragma once
#include <cmath>
#include <array>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../include/lprb200.h"

namespace LPR_381_Group_V22 {

struct InvalidOperationException : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct ArgumentException : std::invalid_argument {
  using std::invalid_argument::invalid_argument;
};

inline void Check(int rc) {
  if (rc != LPR_OK) throw InvalidOperationException(std::string("liblprb200: ") + lpr_last_error());
}

namespace IO {
// InputFileParser.Constraint (IO/InputFileParser.cs:70-82)
struct Constraint {
  std::vector<double> Coefficients;
  std::string Relation;
  double RHS;
};

// IO/InputFileParser.cs:10-68 over the native parser (lpr_model_*): same members, same two silent early returns
// (missing file, fewer than three lines -> Message, members untouched); what throws FormatException /
// IndexOutOfRangeException in the reference throws ArgumentException here, with that name in the text.
class InputFileParser {
 public:
  std::string ProblemType;
  std::vector<double> ObjectiveCoefficients;
  std::vector<Constraint> Constraints;
  std::vector<std::string> SignRestrictions;
  std::string Message;  // the line ReadInputFile writes to the console

  ~InputFileParser() { lpr_model_destroy(model_); }
  InputFileParser() = default;
  InputFileParser(const InputFileParser&) = delete;
  InputFileParser& operator=(const InputFileParser&) = delete;

  void ReadInputFile(const std::string& filePath) {
    lpr_model* m = nullptr;
    if (lpr_model_parse_file(filePath.c_str(), &m) != LPR_OK) throw ArgumentException(lpr_last_error());
    Adopt(m);
  }
  void ReadInputText(const std::string& text) {  // the same on an in-memory buffer
    lpr_model* m = nullptr;
    if (lpr_model_parse_text(text.data(), (int64_t)text.size(), &m) != LPR_OK) throw ArgumentException(lpr_last_error());
    Adopt(m);
  }
  // Program.cs:114-124 and :511-535 applied to the parsed model (Constraints is refreshed)
  void AddCliBoundRows() { Check(lpr_model_add_cli_bound_rows(model_)); PullConstraints(); }
  void AddUpperBoundConstraints() { Check(lpr_model_add_upper_bound_rows(model_)); PullConstraints(); }
  // PrimalSimplexSolver..ctor on the parsed model, built on the device from the native arrays; caller owns it
  lpr_tab* CreateDeviceTableau(bool isMaximization = true, int device = 0) const {
    lpr_tab* t = nullptr;
    Check(lpr_tab_create_from_model(device, model_, isMaximization ? 1 : 0, &t));
    return t;
  }
  const lpr_model* Native() const { return model_; }

 private:
  static std::string Str(int (*fn)(const lpr_model*, char*, int), const lpr_model* m) {
    std::vector<char> buf(1 << 12);
    Check(fn(m, buf.data(), (int)buf.size()));
    return std::string(buf.data());
  }
  void Adopt(lpr_model* m) {
    int loaded = 0, n = 0, rows = 0, ns = 0;
    Check(lpr_model_info(m, &loaded, &n, &rows, &ns));
    Message = Str(lpr_model_message, m);
    if (!loaded) {
      lpr_model_destroy(m);
      return;
    }
    lpr_model_destroy(model_);
    model_ = m;
    ProblemType = Str(lpr_model_problem_type, m);
    std::vector<double> c(n);
    Check(lpr_model_objective(m, c.data()));
    ObjectiveCoefficients.insert(ObjectiveCoefficients.end(), c.begin(), c.end());
    PullConstraints();
    for (int j = 0; j < ns; j++) {
      char buf[1024];
      Check(lpr_model_sign(m, j, buf, (int)sizeof buf));
      SignRestrictions.emplace_back(buf);
    }
  }
  void PullConstraints() {
    int rows = 0;
    Check(lpr_model_info(model_, nullptr, nullptr, &rows, nullptr));
    Constraints.clear();
    for (int i = 0; i < rows; i++) {
      int cnt = 0;
      Constraint c;
      char rel[64];
      Check(lpr_model_constraint(model_, i, nullptr, 0, &cnt, nullptr, 0, nullptr));
      c.Coefficients.resize(cnt);
      Check(lpr_model_constraint(model_, i, c.Coefficients.data(), cnt, nullptr, rel, (int)sizeof rel, &c.RHS));
      c.Relation = rel;
      Constraints.push_back(std::move(c));
    }
  }
  lpr_model* model_ = nullptr;
};

// IO/OutputFileWrite.cs:16-137 over lpr_out_write_*; the parsed model stands in for the reference's four model lists.
// `timestamp` empty = local time now.
struct OutputFileWrite {
  static void WriteFullResults(const std::string& filePath, const std::string& solverUsed, const InputFileParser& model,
                               const std::vector<std::string>& iterationSnapshots, double finalZ,
                               const std::vector<double>* solutionVector, bool append = false,
                               const std::string& timestamp = "") {
    std::vector<const char*> sn;
    for (auto& s : iterationSnapshots) sn.push_back(s.c_str());
    Check(lpr_out_write_full_results(filePath.c_str(), solverUsed.c_str(), model.Native(), sn.data(), (int)sn.size(), finalZ,
                                     solutionVector ? solutionVector->data() : nullptr,
                                     solutionVector ? (int)solutionVector->size() : 0, append ? 1 : 0,
                                     timestamp.empty() ? nullptr : timestamp.c_str()));
  }
  static void WriteSnapshotsOnly(const std::string& filePath, const std::string& solverUsed,
                                 const std::vector<std::string>& snapshots, double finalZ,
                                 const std::vector<double>* solutionVector, bool append = true,
                                 const std::string& timestamp = "") {
    std::vector<const char*> sn;
    for (auto& s : snapshots) sn.push_back(s.c_str());
    Check(lpr_out_write_snapshots_only(filePath.c_str(), solverUsed.c_str(), sn.data(), (int)sn.size(), finalZ,
                                       solutionVector ? solutionVector->data() : nullptr,
                                       solutionVector ? (int)solutionVector->size() : 0, append ? 1 : 0,
                                       timestamp.empty() ? nullptr : timestamp.c_str()));
  }
};
}  // namespace IO

namespace Utilities {
// Utilities/TableIterationFormater.cs:22-48 and NumFormat.N3 (RevisedPrimalSimplexSolver.cs:455-465), native
struct TableIterationFormater {
  static std::string Format(const std::vector<double>& tab, int rows, int cols, int numOriginalVars,
                            const std::string& title, const std::vector<std::string>* rowLabels = nullptr) {
    std::vector<const char*> lab;
    if (rowLabels)
      for (auto& s : *rowLabels) lab.push_back(s.c_str());
    const char* text = nullptr;
    int64_t len = 0;
    Check(lpr_fmt_table(tab.data(), rows, cols, cols, numOriginalVars, title.c_str(), rowLabels ? lab.data() : nullptr,
                        (int)lab.size(), &text, &len));
    return std::string(text, (size_t)len);
  }
  static std::string Format(lpr_tab* deviceTableau, int numOriginalVars, const std::string& title,
                            const std::vector<std::string>* rowLabels = nullptr) {
    std::vector<const char*> lab;
    if (rowLabels)
      for (auto& s : *rowLabels) lab.push_back(s.c_str());
    const char* text = nullptr;
    int64_t len = 0;
    Check(lpr_tab_format(deviceTableau, numOriginalVars, title.c_str(), rowLabels ? lab.data() : nullptr, (int)lab.size(),
                         &text, &len));
    return std::string(text, (size_t)len);
  }
};
struct NumFormat {
  static std::string N3(double x) {
    char buf[400];
    Check(lpr_fmt_n3(x, buf, (int)sizeof buf));
    return buf;
  }
};
}  // namespace Utilities

namespace Simplex {

// Simplex/PrimalSimplexSolver.cs
class PrimalSimplexSolver {
 public:
  double FinalZ = 0.0;
  std::vector<double> SolutionVector;  // empty == null (unbounded / not solved)
  bool HasSolution = false;
  std::vector<double> FinalTableau;    // rows x cols, row-major; empty == null
  int Rows = 0, Cols = 0;
  int Status = LPR_RUNNING;
  std::vector<std::pair<int, int>> PivotLog;  // (row, col)

  PrimalSimplexSolver(const std::vector<double>& objective, const std::vector<IO::Constraint>& constraints,
                      bool isMaximization = true)
      : n_((int)objective.size()), m_((int)constraints.size()) {
    int stride = n_;
    for (auto& c : constraints) stride = std::max<int>(stride, (int)c.Coefficients.size());
    std::vector<double> coef((size_t)m_ * stride, 0.0), rhs(m_);
    std::vector<int> cnt(m_), rel(m_);
    for (int i = 0; i < m_; i++) {
      const auto& c = constraints[i];
      cnt[i] = (int)c.Coefficients.size();
      for (int j = 0; j < cnt[i]; j++) coef[(size_t)i * stride + j] = c.Coefficients[j];
      rel[i] = c.Relation == ">=" ? LPR_REL_GE : (c.Relation == "=" ? LPR_REL_EQ : LPR_REL_LE);
      rhs[i] = c.RHS;
    }
    Check(lpr_tab_create_primal(0, n_, m_, objective.data(), coef.data(), stride, cnt.data(), rel.data(), rhs.data(),
                                isMaximization ? 1 : 0, &tab_));
    Rows = m_ + 1;
    Cols = n_ + m_ + 1;
  }
  ~PrimalSimplexSolver() { lpr_tab_destroy(tab_); }
  PrimalSimplexSolver(const PrimalSimplexSolver&) = delete;
  PrimalSimplexSolver& operator=(const PrimalSimplexSolver&) = delete;

  void Solve() {  // :102-150
    std::vector<int> log(2 * 65536);
    int64_t np = 0;
    Check(lpr_tab_solve(tab_, LPR_RULE_PRIMAL, -1, 0, &Status, &np, log.data(), 65536));
    for (int64_t k = 0; k < np && k < 65536; k++) PivotLog.emplace_back(log[2 * k], log[2 * k + 1]);
    FinalTableau = GetFinalTableau();
    if (Status == LPR_OPTIMAL) {  // :110-126
      Check(lpr_tab_objective(tab_, &FinalZ));
      SolutionVector.assign(n_, 0.0);
      Check(lpr_tab_extract_solution(tab_, n_, SolutionVector.data()));
      HasSolution = true;
    }  // unbounded (:129-135): FinalZ stays 0, SolutionVector stays null
  }
  std::vector<double> GetFinalTableau() const {  // :269-273
    std::vector<double> t((size_t)Rows * Cols);
    Check(lpr_tab_read(tab_, t.data()));
    return t;
  }
  std::vector<int> BasicVariables() const {  // :275-278
    std::vector<int> b(m_);
    Check(lpr_tab_get_basis(tab_, b.data()));
    return b;
  }
  int NumVariables() const { return n_; }

 private:
  int n_, m_;
  lpr_tab* tab_ = nullptr;
};

// Simplex/RevisedPrimalSimplexSolver.cs
class RevisedPrimalSimplexSolver {
 public:
  double FinalZ = 0.0;
  std::vector<double> SolutionVector;
  std::vector<std::array<int, 3>> PivotLog;  // (leaveRow, enter, leaveVar)

  RevisedPrimalSimplexSolver(const std::vector<double>& objective, const std::vector<IO::Constraint>& constraints,
                             bool isMinimization)
      : n_((int)objective.size()), m_((int)constraints.size()) {
    if (objective.empty()) throw ArgumentException("Objective cannot be null or empty.");
    if (constraints.empty()) throw ArgumentException("Constraints cannot be null or empty.");
    std::vector<double> A((size_t)m_ * n_), b(m_);
    for (int i = 0; i < m_; i++) {
      if ((int)constraints[i].Coefficients.size() != n_)
        throw ArgumentException("Constraint " + std::to_string(i + 1) + " has incorrect number of coefficients.");
      for (int j = 0; j < n_; j++) A[(size_t)i * n_ + j] = constraints[i].Coefficients[j];
      b[i] = constraints[i].RHS;  // Relation ignored (:55-61)
    }
    Check(lpr_rev_create(0, m_, n_, A.data(), b.data(), objective.data(), isMinimization ? 1 : 0, &rev_));
  }
  ~RevisedPrimalSimplexSolver() { lpr_rev_destroy(rev_); }
  RevisedPrimalSimplexSolver(const RevisedPrimalSimplexSolver&) = delete;
  RevisedPrimalSimplexSolver& operator=(const RevisedPrimalSimplexSolver&) = delete;

  std::vector<std::string> IterationSnapshots;  // one CaptureSnapshot block per iteration + "Optimal" (:226-246, :124-146)
  // below this many table elements Solve() steps iteration by iteration and records the reference's text blocks
  // (SURVEY 8b "Snapshots"); above it the loop runs entirely on the device and no text is built
  long TraceMaxElements = 4096;

  void Solve() {  // :82-251; exceptions :91, :179, :267
    int st = 0;
    if ((long)m_ * (n_ + m_ + 1) <= TraceMaxElements) {
      Check(lpr_rev_begin(rev_));
      while (true) {
        int e = -1, lr = -1, lv = -1;
        Check(lpr_rev_step(rev_, &st, &e, &lr, &lv));
        if (st != LPR_RUNNING && st != LPR_OPTIMAL) break;
        const char* text = nullptr;
        int64_t len = 0;
        Check(lpr_rev_format_snapshot(rev_, &text, &len));
        IterationSnapshots.emplace_back(text, (size_t)len);
        if (st == LPR_OPTIMAL) break;
        PivotLog.push_back({lr, e, lv});
      }
    } else {
      int64_t nit = 0;
      std::vector<int> log(3 * 65536);
      Check(lpr_rev_solve(rev_, -1, 0, &st, &nit, log.data(), 65536));
      for (int64_t k = 0; k < nit && k < 65536; k++) PivotLog.push_back({log[3 * k], log[3 * k + 1], log[3 * k + 2]});
    }
    if (st == LPR_INFEASIBLE) throw std::runtime_error("Infeasible basis (negative basic value).");
    if (st == LPR_UNBOUNDED) throw std::runtime_error("Unbounded problem (no positive component in direction).");
    if (st == LPR_PIVOT_TOO_SMALL) throw std::runtime_error("Pivot too small.");
    SolutionVector.assign(n_, 0.0);
    Check(lpr_rev_read_x(rev_, SolutionVector.data()));
    Check(lpr_rev_read_z(rev_, &FinalZ));
  }
  std::vector<int> BasicVariables() const {
    std::vector<int> b(m_);
    Check(lpr_rev_read_basis(rev_, b.data()));
    return b;
  }
  std::vector<double> DualPrices() const {  // y = c_B B^-1 (:93)
    std::vector<double> y(m_);
    Check(lpr_rev_read_y(rev_, y.data()));
    return y;
  }

 private:
  int n_, m_;
  lpr_rev* rev_ = nullptr;
};

// shared by PrimalSimplexSolver2 / DualSimplexSolver: (objectiveRow, constraintRows) <-> device tableau
inline lpr_tab* UploadRows(const std::vector<double>& obj, const std::vector<std::vector<double>>& rows, int headroom) {
  if (rows.empty()) throw ArgumentException("No constraint rows.");
  const int w = (int)obj.size();
  for (auto& r : rows)
    if ((int)r.size() != w) throw ArgumentException("All rows (obj & constraints) must have the same length.");
  std::vector<double> T;
  T.insert(T.end(), obj.begin(), obj.end());
  for (auto& r : rows) T.insert(T.end(), r.begin(), r.end());
  lpr_tab* h = nullptr;
  Check(lpr_tab_create(0, (int)rows.size() + 1, w, (int)rows.size() + 1 + headroom, w, T.data(), &h));
  return h;
}
inline void DownloadRows(lpr_tab* h, std::vector<double>& obj, std::vector<std::vector<double>>& rows) {
  int R = 0, C = 0, ld = 0;
  Check(lpr_tab_dims(h, &R, &C, &ld));
  std::vector<double> T((size_t)R * C);
  Check(lpr_tab_read(h, T.data()));
  obj.assign(T.begin(), T.begin() + C);
  rows.resize(R - 1);
  for (int i = 1; i < R; i++) rows[i - 1].assign(T.begin() + (size_t)i * C, T.begin() + (size_t)(i + 1) * C);
}

// Simplex/PrimalSimplexSolver2.cs
class PrimalSimplexSolver2 {
 public:
  double FinalZ = 0.0;
  PrimalSimplexSolver2(const std::vector<double>& objectiveRow, const std::vector<std::vector<double>>& constraintRows)
      : tab_(UploadRows(objectiveRow, constraintRows, 0)) {}
  ~PrimalSimplexSolver2() { lpr_tab_destroy(tab_); }
  bool Solve(int maxIters = 10000, bool printSteps = false) {  // :46-97
    int st = 0;
    int64_t np = 0;
    Check(lpr_tab_solve(tab_, LPR_RULE_PRIMAL2, maxIters, printSteps ? 1 : 0, &st, &np, nullptr, 0));
    if (st == LPR_PIVOT_TOO_SMALL) throw InvalidOperationException("Pivot too small/zero.");
    optimal_ = (st == LPR_OPTIMAL);
    if (optimal_) Check(lpr_tab_objective(tab_, &FinalZ));
    return optimal_;
  }
  void GetRows(std::vector<double>& obj, std::vector<std::vector<double>>& rows, bool solveIfNeeded = true) {
    if (!optimal_ && solveIfNeeded && !Solve())
      throw InvalidOperationException("Could not reach an optimal tableau (unbounded or infeasible).");
    DownloadRows(tab_, obj, rows);
  }

 private:
  lpr_tab* tab_;
  bool optimal_ = false;
};

// Simplex/DualSimplex.cs (global namespace in the reference)
class DualSimplexSolver {
 public:
  // mutates objectiveRow / constraintRows in place, like the reference
  bool Solve(std::vector<double>& objectiveRow, std::vector<std::vector<double>>& constraintRows, int maxIters = 10000,
             bool printSteps = true) {
    lpr_tab* h = UploadRows(objectiveRow, constraintRows, 0);
    int st = 0;
    int64_t np = 0;
    int rc = lpr_tab_solve(h, LPR_RULE_DUAL, maxIters, printSteps ? 1 : 0, &st, &np, nullptr, 0);
    if (rc == LPR_OK) DownloadRows(h, objectiveRow, constraintRows);
    lpr_tab_destroy(h);
    Check(rc);
    if (st == LPR_PIVOT_TOO_SMALL) throw InvalidOperationException("Pivot too small/zero.");
    return st == LPR_OPTIMAL;
  }
  static bool AnyNegativeRhs(const std::vector<std::vector<double>>& rows) {  // :180-185
    for (auto& r : rows)
      if (!r.empty() && r.back() < -1e-9) return true;
    return false;
  }
};

}  // namespace Simplex

namespace IntegerProgramming {

// IntegerProgramming/CuttingPlaneSolver.cs:64-229 (in place; cut rows are appended)
class CuttingPlaneSolver {
 public:
  int Status = LPR_RUNNING;
  void CuttingPlaneSolution(std::vector<double>& objectiveRow, std::vector<std::vector<double>>& constraintRows,
                            int maxCuts = -1) {
    lpr_tab* h = Simplex::UploadRows(objectiveRow, constraintRows, maxCuts < 0 ? 64 : maxCuts + 1);
    int ncuts = 0;
    int rc = lpr_tab_cutting_plane(h, maxCuts, &Status, &ncuts, nullptr, 0);
    if (rc == LPR_OK) Simplex::DownloadRows(h, objectiveRow, constraintRows);
    lpr_tab_destroy(h);
    Check(rc);
  }
};

// IntegerProgramming/BranchAndBoundAdapter.cs:9-24
struct BranchAndBoundAdapter {
  static std::pair<std::vector<double>, double> SolveFromPrimal(const Simplex::PrimalSimplexSolver& primal,
                                                                bool enablePruning = false, bool isMin = false,
                                                                int64_t maxNodes = 20 /* reference cap :1038 */) {
    (void)isMin;  // ignored by the reference as well
    if (primal.FinalTableau.empty()) throw InvalidOperationException("Primal simplex has not been solved yet.");
    const int n = primal.HasSolution ? (int)primal.SolutionVector.size() : std::max(1, primal.Cols - 1);
    std::vector<double> x(n);
    double z = 0;
    int has = 0, st = 0;
    int64_t nodes = 0, piv = 0;
    Check(lpr_bb_solve(0, primal.Rows, primal.Cols, primal.FinalTableau.data(), n, enablePruning ? 1 : 0, maxNodes,
                       x.data(), &z, &has, &nodes, &piv, nullptr, nullptr, 0, &st));
    if (!has) return {std::vector<double>(), -std::numeric_limits<double>::infinity()};
    return {x, z};
  }
};

// IntegerProgramming/BranchBoundSimplexSolver.cs: the RunBranchAndBound entry (:28-113, :281-468, :1233-1298).
// Model rows are the reference's ragged [coefficients..., rhs, type flag] lists.
struct BranchBoundSimplexSolver {
  class DualSimplexSolverBB {
   public:
    // FormulateTableau :28-113 built on the device; returns the (m+1) x (n+m+1) tableau row-major and, like the
    // reference, negates the '>=' rows of the caller's list and drops the flags
    std::vector<double> FormulateTableau(const std::vector<double>& objectiveFunction,
                                         std::vector<std::vector<double>>& constraints, int* rows = nullptr,
                                         int* cols = nullptr) {
      lpr_tab* h = Create(objectiveFunction, constraints);
      int R = 0, C = 0, ld = 0;
      lpr_tab_dims(h, &R, &C, &ld);
      std::vector<double> T((size_t)R * C);
      int rc = lpr_tab_read(h, T.data());
      lpr_tab_destroy(h);
      Check(rc);
      StripFlags(constraints);
      if (rows) *rows = R;
      if (cols) *cols = C;
      return T;
    }
    // DoDualSimplex :289-468 without tableauOverride; returns false when the reference returns optimalValue = null
    bool DoDualSimplex(const std::vector<double>& objectiveFunction, std::vector<std::vector<double>>& constraints,
                       bool isMinimization, std::vector<double>& finalTableau, int& rows, int& cols,
                       double& optimalValue, bool round4 = false) {
      lpr_tab* h = Create(objectiveFunction, constraints);
      StripFlags(constraints);
      int st = 0, ld = 0;
      int64_t npiv = 0;
      int rc = lpr_tab_bb_node_solve_ex(h, isMinimization ? 1 : 0, -1, &st, &npiv, nullptr, 0);
      if (rc == LPR_OK && round4) rc = lpr_tab_round4(h);  // RoundAllTableaux :540-593
      if (rc == LPR_OK) rc = lpr_tab_dims(h, &rows, &cols, &ld);
      if (rc == LPR_OK) {
        finalTableau.assign((size_t)rows * cols, 0.0);
        rc = lpr_tab_read(h, finalTableau.data());
      }
      lpr_tab_destroy(h);
      Check(rc);
      if (st != LPR_OPTIMAL) return false;
      optimalValue = finalTableau[(size_t)cols - 1];
      return true;
    }

   private:
    static lpr_tab* Create(const std::vector<double>& obj, const std::vector<std::vector<double>>& cons) {
      const int n = (int)obj.size(), m = (int)cons.size();
      int stride = 2;
      for (auto& c : cons) stride = std::max<int>(stride, (int)c.size());
      std::vector<double> flat((size_t)std::max(1, m) * stride, 0.0);
      std::vector<int> len(std::max(1, m), 0);
      for (int i = 0; i < m; i++) {
        std::copy(cons[i].begin(), cons[i].end(), flat.begin() + (size_t)i * stride);
        len[i] = (int)cons[i].size();
      }
      lpr_tab* h = nullptr;
      Check(lpr_tab_create_bb(0, n, m, obj.data(), flat.data(), stride, len.data(), 0, 0, &h));
      return h;
    }
    static void StripFlags(std::vector<std::vector<double>>& cons) {  // :42-56
      for (auto& row : cons)
        if (!row.empty() && row.back() == 1)
          for (auto& v : row) v = -1 * v;
      for (auto& row : cons)
        if (!row.empty()) row.pop_back();
    }
  };

  class BranchAndBound {
   public:
    std::vector<double> objectiveCoefficients;
    static void ConfigureProblem(const std::vector<double>& objective, std::vector<std::vector<double>>& constraints) {
      const int n = (int)objective.size();  // :1233-1251: rows one entry longer than a model row
      for (int i = 0; i < n; i++) {
        std::vector<double> row(n + 3, 0.0);
        row[i] = 1;
        row[n + 1] = 1;
        constraints.push_back(row);
      }
    }
    // :1253-1298; returns (bestSolution, bestValue); an empty solution == the reference's null
    std::pair<std::vector<double>, double> RunBranchAndBound(const std::vector<double>& objectivePassed,
                                                             const std::vector<std::vector<double>>& constraintsPassed,
                                                             bool isMin, int64_t maxNodes = 20) {
      objectiveCoefficients = objectivePassed;
      auto cons = constraintsPassed;
      ConfigureProblem(objectiveCoefficients, cons);
      DualSimplexSolverBB solver;
      std::vector<double> T;
      int R = 0, C = 0;
      double opt = 0;
      if (!solver.DoDualSimplex(objectiveCoefficients, cons, isMin, T, R, C, opt, /*round4=*/true))
        throw InvalidOperationException("initial LP relaxation failed");
      const int n = (int)objectiveCoefficients.size();
      std::vector<double> x(n);
      double z = 0;
      int has = 0, st = 0;
      int64_t nodes = 0, piv = 0;
      Check(lpr_bb_solve(0, R, C, T.data(), n, 0, maxNodes, x.data(), &z, &has, &nodes, &piv, nullptr, nullptr, 0, &st));
      if (!has) return {std::vector<double>(), -std::numeric_limits<double>::infinity()};
      return {x, z};
    }
  };
};

// Program.cs:444-463 (the class is missing from the reference; contract from the call site)
struct KnapsackItem {
  int Id;
  double Value, Weight;
};
class KnapsackBranchBoundSimplex {
 public:
  KnapsackBranchBoundSimplex(int capacity, std::vector<double> weights, std::vector<double> values)
      : cap_(capacity), w_(std::move(weights)), v_(std::move(values)), chosen_(w_.size(), 0) {}
  int Gpus = 1;  // > 1: the open-node pool is partitioned over that many GPUs inside the library (lpr_knap_solve_mgpu)
  double Solve() {
    int st = 0;
    if (Gpus > 1)
      Check(lpr_knap_solve_mgpu(Gpus, nullptr, cap_, (int)w_.size(), w_.data(), v_.data(), -1, -1, 0.0, &best_,
                                chosen_.data(), &nodes_, &st, nullptr));
    else
      Check(lpr_knap_solve(0, cap_, (int)w_.size(), w_.data(), v_.data(), -1, &best_, chosen_.data(), &nodes_, &st));
    return best_;
  }
  std::vector<KnapsackItem> GetSelectedItemsOriginal() const {
    std::vector<KnapsackItem> out;
    for (size_t i = 0; i < w_.size(); i++)
      if (chosen_[i]) out.push_back({(int)i, v_[i], w_[i]});
    return out;
  }
  int64_t Nodes() const { return nodes_; }

 private:
  double cap_;
  std::vector<double> w_, v_;
  std::vector<uint8_t> chosen_;
  double best_ = 0;
  int64_t nodes_ = 0;
};
struct KnapsackBranchBoundSolver {  // Program.cs:468 DP arbiter
  static double Solve(int capacity, const std::vector<int>& weights, const std::vector<int>& values) {
    double best = 0;
    std::vector<uint8_t> ch(weights.size());
    Check(lpr_knap_dp(0, capacity, (int)weights.size(), weights.data(), values.data(), &best, ch.data()));
    return best;
  }
};

}  // namespace IntegerProgramming

namespace SensitivityAnalysis {

// SensitivityAnalysis/SensitivityAnalyzer.cs: ctor :22-41, ResolveAll :203-209 (RebuildBasicsFromTableau :706-723,
// DualSimplexIfNeeded :168-201, ReOptimize :121-166), AddNewConstraintNonInteractive :609-659, ShadowPrices
// :212-222.  The tableau stays on the device between calls.
class SensitivityAnalyzer {
 public:
  SensitivityAnalyzer(const std::vector<double>& finalTableau, int rows, int cols, const std::vector<double>& solution,
                      double zValue, const std::vector<int>& basicVariables, int headroom = 16)
      : solutionVector_(solution), finalZ_(zValue) {
    std::vector<double> T(finalTableau);
    T[(size_t)cols - 1] = zValue;  // :33
    Check(lpr_tab_create(0, rows, cols, rows + headroom, cols + headroom, T.data(), &h_));
    if (!basicVariables.empty()) lpr_tab_set_basis(h_, basicVariables.data());
    Check(lpr_tab_sens_rebuild_basis(h_));  // :36
  }
  ~SensitivityAnalyzer() { lpr_tab_destroy(h_); }
  SensitivityAnalyzer(const SensitivityAnalyzer&) = delete;
  SensitivityAnalyzer& operator=(const SensitivityAnalyzer&) = delete;

  double CurrentZ() const { return finalZ_; }
  const std::vector<double>& SolutionVector() const { return solutionVector_; }
  std::vector<double> CurrentTableau(int* rows = nullptr, int* cols = nullptr) const {
    int R = 0, C = 0, ld = 0;
    Check(lpr_tab_dims(h_, &R, &C, &ld));
    std::vector<double> T((size_t)R * C);
    Check(lpr_tab_read(h_, T.data()));
    if (rows) *rows = R;
    if (cols) *cols = C;
    return T;
  }
  std::vector<int> BasicVariables() const {
    int R = 0, C = 0, ld = 0;
    Check(lpr_tab_dims(h_, &R, &C, &ld));
    std::vector<int> b(std::max(1, R - 1));
    Check(lpr_tab_get_basis(h_, b.data()));
    b.resize(R - 1);
    return b;
  }
  std::vector<double> ShadowPrices() const {  // :212-222
    int R = 0, C = 0, ld = 0;
    Check(lpr_tab_dims(h_, &R, &C, &ld));
    std::vector<double> row0(C);
    Check(lpr_tab_read_row(h_, 0, row0.data()));
    const int m = R - 1, n = C - m - 1;
    return std::vector<double>(row0.begin() + n, row0.begin() + n + m);
  }
  void ResolveAll(int maxIter = 10000) {  // :203-209
    Check(lpr_tab_sens_rebuild_basis(h_));
    int st = 0;
    int64_t npiv = 0;
    Check(lpr_tab_solve(h_, LPR_RULE_SENS, maxIter, 0, &st, &npiv, nullptr, 0));
    if (st == LPR_INFEASIBLE) throw InvalidOperationException("Infeasible after RHS change (dual simplex).");
    if (st == LPR_UNBOUNDED) throw InvalidOperationException("Unbounded during re-optimization.");
    if (st == LPR_ITER_LIMIT) throw InvalidOperationException("Re-optimization exceeded iteration limit.");
    Check(lpr_tab_objective(h_, &finalZ_));
    int R = 0, C = 0, ld = 0;
    Check(lpr_tab_dims(h_, &R, &C, &ld));
    solutionVector_.assign(C - 1, 0.0);
    Check(lpr_tab_sens_solution(h_, solutionVector_.data()));
  }
  void AddNewConstraintNonInteractive(const std::vector<double>& tech, double rhs) {  // :609-659
    int R = 0, C = 0, ld = 0;
    Check(lpr_tab_dims(h_, &R, &C, &ld));
    if ((int)tech.size() < C - 1) throw ArgumentException("tech needs one coefficient per tableau column");
    double aX = 0.0;
    for (size_t j = 0; j < std::min(tech.size(), solutionVector_.size()); j++) aX += tech[j] * solutionVector_[j];
    Check(lpr_tab_sens_add_constraint(h_, tech.data(), rhs - aX));
    ResolveAll();
  }

 private:
  lpr_tab* h_ = nullptr;
  std::vector<double> solutionVector_;
  double finalZ_;
};

}  // namespace SensitivityAnalysis
}  // namespace LPR_381_Group_V22
