// ivp_oracle_implicit.hpp -- RADAU / BDF restatement (placeholder until the implicit path lands).
// TEST INFRASTRUCTURE ONLY (see ivp_oracle.hpp header).
#pragma once
namespace oracle {
namespace radau {
inline void interpolate(double, double*, size_t, const double*, double, double) { throw ConfigError("RADAU oracle not built yet"); }
template <class F, class S>
IntegrationResult solve(const F&, double, const std::vector<double>&, double, const Tol&, const Tol&, const StepCfg&, S*) {
  throw ConfigError("RADAU oracle not built yet");
}
}  // namespace radau
namespace bdf {
inline void interpolate(double, double*, size_t, const double*, double, double) { throw ConfigError("BDF oracle not built yet"); }
template <class F, class S>
IntegrationResult solve(const F&, double, const std::vector<double>&, double, const Tol&, const Tol&, const StepCfg&, S*) {
  throw ConfigError("BDF oracle not built yet");
}
}  // namespace bdf
}  // namespace oracle
