"""Model types at the boundary: IO/InputFileParser.cs (parser :19-68, Constraint :70-82)."""
import os


class Constraint:
    """InputFileParser.Constraint (IO/InputFileParser.cs:70-82)."""

    def __init__(self, coefficients, relation, rhs):
        self.Coefficients = list(coefficients)
        self.Relation = relation
        self.RHS = float(rhs)

    def __repr__(self):
        return f"Constraint({self.Coefficients}, {self.Relation!r}, {self.RHS})"


class InputFileParser:
    """IO/InputFileParser.cs:10-68: `max|min c...` / `a... rel rhs` / sign line."""

    def __init__(self):
        self.ProblemType = None
        self.ObjectiveCoefficients = []
        self.Constraints = []
        self.SignRestrictions = []

    def ReadInputFile(self, file_path):
        if not os.path.exists(file_path):
            print("Sorry, we can't find your file, please check it's in the right folser")
            return
        with open(file_path, "r", encoding="utf-8-sig") as f:
            lines = f.read().splitlines()
        if len(lines) < 3:
            print("The input file is not formatted correctly.")
            return
        objective_line = lines[0].strip().split(" ")
        self.ProblemType = objective_line[0].lower()
        for tok in objective_line[1:]:
            self.ObjectiveCoefficients.append(float(tok))
        n = len(self.ObjectiveCoefficients)
        for line in lines[1:-1]:
            parts = [p for p in line.strip().split(" ") if p]
            coeffs = [float(parts[j]) for j in range(n)]
            self.Constraints.append(Constraint(coeffs, parts[n], float(parts[n + 1])))
        self.SignRestrictions.extend(lines[-1].strip().split(" "))
        print("Your file was read and is in the correct format!")


def add_cli_bound_rows(n, constraints):
    """Program.cs:114-124 / :372-382: menu options 1 and 3 append `x_i <= 1` for every variable;
    the coefficient list has length n+3 with a stray 1 at index n+1."""
    for i in range(n):
        co = [0.0] * (n + 3)
        co[i] = 1.0
        co[n + 1] = 1.0
        constraints.append(Constraint(co, "<=", 1.0))
    return constraints


def add_upper_bound_constraints(n, sign_restrictions, constraints):
    """Program.cs:511-535 AddUpperBoundConstraints (menu option 2)."""
    if not sign_restrictions:
        return constraints
    for j in range(n):
        sr = sign_restrictions[min(j, len(sign_restrictions) - 1)] or ""
        s = sr.replace(" ", "")
        if "bin" in s.lower() or "≤1" in s or "<=1" in s:
            co = [0.0] * n
            co[j] = 1.0
            constraints.append(Constraint(co, "<=", 1.0))
    return constraints
