// ORACLE (test infrastructure only) -- the synthetic problems the reference's own solver tests run on.
//   linear system + quadratic costs, optional jump map / event cost / mode-switched constraint:
//       ocs2_oc/test/include/ocs2_oc/test/testProblemsGeneration.h:45-101 (getRandomCost/Dynamics/Constraints),
//       ocs2_sqp/ocs2_sqp/test/testUnconstrained.cpp:40-93, testSwitchedProblem.cpp:47-153
//   circular kinematics: ocs2_oc/test/include/ocs2_oc/test/circular_kinematics.h:47-136
#pragma once
#include "sqp.hpp"

namespace orc {

// x' = A x + B u ; L = 1/2 dx'Q dx + 1/2 du'R du + du'P dx (ocs2::QuadraticStateInputCost,
// ocs2_core/src/cost/QuadraticStateInputCost.cpp:71-99) about (xRef,uRef); optional per-node constraint rows.
struct LinearQuadraticOcp : Ocp {
  Mat A, B, G;                  // flow map, jump map
  Mat Q, R, P, Qf, Qe;          // intermediate (twice: cost + softConstraint handled by caller summing), final, event
  Vec xRef, uRef;
  bool hasEventCost = false;
  std::vector<int> nodeMode;    // per node: -1 no constraint, else index into Cm/Dm/em
  std::vector<Mat> Cm, Dm;
  std::vector<Vec> em;

  Vec flowMap(int, double, const Vec& x, const Vec& u) override { return mul(A, x) + mul(B, u); }
  LinApprox flowMapLin(int k, double t, const Vec& x, const Vec& u) override {
    LinApprox l;
    l.f = flowMap(k, t, x, u);
    l.dfdx = A;
    l.dfdu = B;
    return l;
  }
  double cost(int, double, const Vec& x, const Vec& u) override {
    const Vec dx = x - xRef, du = u - uRef;
    return 0.5 * dot(dx, mul(Q, dx)) + 0.5 * dot(du, mul(R, du)) + dot(du, mul(P, dx));
  }
  QuadApprox costQuad(int k, double t, const Vec& x, const Vec& u) override {
    const Vec dx = x - xRef, du = u - uRef;
    QuadApprox c;
    c.f = cost(k, t, x, u);
    c.dfdxx = Q;
    c.dfduu = R;
    c.dfdux = P;
    c.dfdx = mul(Q, dx) + mul(P, du, true);
    c.dfdu = mul(R, du) + mul(P, dx);
    return c;
  }
  double finalCost(int, double, const Vec& x) override {
    const Vec dx = x - xRef;
    return 0.5 * dot(dx, mul(Qf, dx));
  }
  QuadApprox finalCostQuad(int k, double t, const Vec& x) override {
    QuadApprox c;
    c.f = finalCost(k, t, x);
    c.dfdxx = Qf;
    c.dfdx = mul(Qf, x - xRef);
    return c;
  }
  Vec jumpMap(int, double, const Vec& x) override { return G.r ? mul(G, x) : x; }
  Mat jumpMapDx(int, double, const Vec&) override { return G.r ? G : Mat::identity(nx); }
  double eventCost(int, double, const Vec& x) override {
    if (!hasEventCost) return 0.0;
    const Vec dx = x - xRef;
    return 0.5 * dot(dx, mul(Qe, dx));
  }
  QuadApprox eventCostQuad(int k, double t, const Vec& x) override {
    QuadApprox c;
    c.dfdxx = Mat(nx, nx);
    c.dfdx = vzero(nx);
    if (hasEventCost) {
      c.f = eventCost(k, t, x);
      c.dfdxx = Qe;
      c.dfdx = mul(Qe, x - xRef);
    }
    return c;
  }
  Vec eqConstraint(int k, double, const Vec& x, const Vec& u) override {
    const int m = nodeMode.empty() ? -1 : nodeMode[k];
    if (m < 0) return {};
    return mul(Cm[m], x) + mul(Dm[m], u) + em[m];
  }
  LinApprox eqConstraintLin(int k, double t, const Vec& x, const Vec& u) override {
    const int m = nodeMode.empty() ? -1 : nodeMode[k];
    if (m < 0) return {};
    LinApprox l;
    l.f = eqConstraint(k, t, x, u);
    l.dfdx = Cm[m];
    l.dfdu = Dm[m];
    return l;
  }
};

struct CircularKinematicsOcp : Ocp {
  CircularKinematicsOcp() {
    nx = 2;
    nu = 2;
  }
  Vec flowMap(int, double, const Vec&, const Vec& u) override { return u; }
  LinApprox flowMapLin(int, double, const Vec&, const Vec& u) override {
    LinApprox l;
    l.f = u;
    l.dfdx = Mat(2, 2);
    l.dfdu = Mat::identity(2);
    return l;
  }
  double cost(int, double, const Vec& x, const Vec& u) override {
    const double g = x[0] * u[1] - x[1] * u[0] - 1.0;
    return 0.5 * g * g + 0.005 * dot(u, u);
  }
  QuadApprox costQuad(int k, double t, const Vec& x, const Vec& u) override {
    // exact Hessian (the reference differentiates the scalar cost with CppAD, StateInputCostCppAd)
    const double g = x[0] * u[1] - x[1] * u[0] - 1.0;
    const double gx[2] = {u[1], -u[0]}, gu[2] = {-x[1], x[0]};
    QuadApprox c;
    c.f = cost(k, t, x, u);
    c.dfdx = {g * gx[0], g * gx[1]};
    c.dfdu = {g * gu[0] + 0.01 * u[0], g * gu[1] + 0.01 * u[1]};
    c.dfdxx = Mat(2, 2);
    c.dfduu = Mat(2, 2);
    c.dfdux = Mat(2, 2);
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) {
        c.dfdxx(i, j) = gx[i] * gx[j];
        c.dfduu(i, j) = gu[i] * gu[j] + (i == j ? 0.01 : 0.0);
        c.dfdux(i, j) = gu[i] * gx[j];
      }
    // g * d2g: d2g/dx0du1 = 1, d2g/dx1du0 = -1
    c.dfdux(1, 0) += g;
    c.dfdux(0, 1) -= g;
    return c;
  }
  double finalCost(int, double, const Vec&) override { return 0.0; }
  QuadApprox finalCostQuad(int, double, const Vec&) override {
    QuadApprox c;
    c.dfdxx = Mat(2, 2);
    c.dfdx = vzero(2);
    return c;
  }
  Vec eqConstraint(int, double, const Vec& x, const Vec& u) override { return {dot(x, u)}; }
  LinApprox eqConstraintLin(int, double, const Vec& x, const Vec& u) override {
    LinApprox l;
    l.f = {dot(x, u)};
    l.dfdx = Mat(1, 2);
    l.dfdu = Mat(1, 2);
    for (int j = 0; j < 2; ++j) {
      l.dfdx(0, j) = u[j];
      l.dfdu(0, j) = x[j];
    }
    return l;
  }
};

}  // namespace orc
