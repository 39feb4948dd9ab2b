// bppgpu shim (see ../bppgpu_shim.hpp): batched Brent line searches, ChromosomeNumberOptimizer, PseudoNewtonOptimizer
#pragma once
#include "ancestral.hpp"

namespace bppshim {

// ---- batched line searches for the multi-start optimiser ------------------------------------------------------------
// ChromosomeNumberOptimizer::optimizeModelParametersOneDimension (Likelihood/ChromosomeNumberOptimizer.cpp:440-530) runs, for every
// starting point in turn and every parameter in turn, a bounded Brent search whose every probe is a full likelihood evaluation.
// The probes of DIFFERENT points are independent, so here all points search the same parameter in lockstep: one probe per point
// per step, ALL of them evaluated by one device call.  The search itself is Brent's bounded minimiser (golden section + parabolic
// interpolation on [lower, upper]; Forsythe, Malcolm & Moler's fmin) written as a per-point state machine; bpp-core's
// BrentOneDimension is not in the reference tree, so probe sequences are not claimed to match it -- optima are.
//
// F: size(), setParameter(point, name, value), evaluate() [one batched evaluation], value(point), parameter(point, name).
template <class F>
class BatchedBrent {
 public:
  explicit BatchedBrent(F* f) : f_(f), nbBatchEvaluations_(0) {}
  // minimise over `name` in [lower, upper] for every point with active[k]; a point keeps its current value if the search found
  // nothing better.  Returns the number of batched evaluations.
  unsigned search(const std::string& name, double lower, double upper, double tol, const std::vector<char>& active, unsigned maxSteps = 200) {
    const size_t K = f_->size();
    const double c = 0.5 * (3.0 - std::sqrt(5.0)), eps = std::sqrt(std::numeric_limits<double>::epsilon());
    std::vector<State> st(K);
    unsigned evals = 0;
    for (size_t k = 0; k < K; ++k) {
      State& s = st[k];
      s.done = !active[k];
      s.x0 = f_->parameter(k, name);
      s.f0 = f_->value(k);
      s.a = lower; s.b = upper;
      s.v = s.w = s.x = s.a + c * (s.b - s.a);
      s.e = s.d = 0;
      s.u = s.x;
      if (!s.done) f_->setParameter(k, name, s.u);
    }
    f_->evaluate(); ++evals;
    for (size_t k = 0; k < K; ++k) if (!st[k].done) st[k].fx = st[k].fv = st[k].fw = f_->value(k);
    for (unsigned step = 0; step < maxSteps; ++step) {
      bool any = false;
      for (size_t k = 0; k < K; ++k) {
        State& s = st[k];
        if (s.done) continue;
        const double xm = 0.5 * (s.a + s.b), tol1 = eps * std::fabs(s.x) + tol / 3.0, tol2 = 2.0 * tol1;
        if (std::fabs(s.x - xm) <= tol2 - 0.5 * (s.b - s.a)) { s.done = true; continue; }
        bool golden = true;
        if (std::fabs(s.e) > tol1) {   // parabolic step through x, w, v
          double r = (s.x - s.w) * (s.fx - s.fv), q = (s.x - s.v) * (s.fx - s.fw), pp = (s.x - s.v) * q - (s.x - s.w) * r;
          q = 2.0 * (q - r);
          if (q > 0) pp = -pp;
          q = std::fabs(q);
          r = s.e;
          s.e = s.d;
          if (std::fabs(pp) < std::fabs(0.5 * q * r) && pp > q * (s.a - s.x) && pp < q * (s.b - s.x)) {
            s.d = pp / q;
            const double u = s.x + s.d;
            if (u - s.a < tol2 || s.b - u < tol2) s.d = xm >= s.x ? tol1 : -tol1;
            golden = false;
          }
        }
        if (golden) {
          s.e = s.x >= xm ? s.a - s.x : s.b - s.x;
          s.d = c * s.e;
        }
        s.u = std::fabs(s.d) >= tol1 ? s.x + s.d : s.x + (s.d >= 0 ? tol1 : -tol1);
        f_->setParameter(k, name, s.u);
        any = true;
      }
      if (!any) break;
      f_->evaluate(); ++evals;
      for (size_t k = 0; k < K; ++k) {
        State& s = st[k];
        if (s.done) continue;
        const double fu = f_->value(k);
        if (fu <= s.fx) {
          if (s.u >= s.x) s.a = s.x; else s.b = s.x;
          s.v = s.w; s.fv = s.fw; s.w = s.x; s.fw = s.fx; s.x = s.u; s.fx = fu;
        } else {
          if (s.u < s.x) s.a = s.u; else s.b = s.u;
          if (fu <= s.fw || s.w == s.x) { s.v = s.w; s.fv = s.fw; s.w = s.u; s.fw = fu; }
          else if (fu <= s.fv || s.v == s.x || s.v == s.w) { s.v = s.u; s.fv = fu; }
        }
      }
    }
    // settle every searched point on the better of (start, best found); one more batched evaluation brings values() in line
    for (size_t k = 0; k < K; ++k)
      if (active[k]) f_->setParameter(k, name, st[k].fx < st[k].f0 ? st[k].x : st[k].x0);
    f_->evaluate(); ++evals;
    nbBatchEvaluations_ += evals;
    return evals;
  }
  unsigned getNumberOfBatchEvaluations() const { return nbBatchEvaluations_; }

 private:
  struct State {
    bool done;
    double a, b, x, w, v, fx, fw, fv, d, e, u, x0, f0;
  };
  F* f_;
  unsigned nbBatchEvaluations_;
};

// The multi-start driver of ChromosomeNumberOptimizer::optimize (Likelihood/ChromosomeNumberOptimizer.cpp:115-153) on one device
// object: cycle i keeps the numOfPoints[i] best points (the reference sorts its vector of likelihoods, :150) and runs
// numOfIterations[i] rounds of per-parameter line searches (optimizeModelParametersOneDimension), stopping a round early when no
// point improved by more than the tolerance (:520-522, applied to the best change over the points).  Parameters are searched on
// (lowerBoundOfRateParam + 1e-10, upperBoundOfRateParam] = (0, 100] (ChromosomeSubstitutionModel.h:15-18; :505).
class ChromosomeNumberOptimizer {
 public:
  ChromosomeNumberOptimizer(const Tree& tree, const VectorSiteContainer& data, const std::vector<ChromosomeSubstitutionModel*>& startPoints,
                            DiscreteDistribution* rDist, bool weightedRootFreq = true, int device = 0)
      : models_(startPoints), batch_(tree, data, weightedRootFreq, std::vector<SubstitutionModel*>(startPoints.begin(), startPoints.end()), rDist, device),
        adaptor_(this), brent_(&adaptor_), active_(startPoints.size(), 1), order_(startPoints.size()) {
    for (size_t k = 0; k < order_.size(); ++k) order_[k] = k;
    batch_.initialize();
  }
  void setParameterNames(const std::vector<std::string>& names) { paramNames_ = names; }
  // numOfPoints / numOfIterations as in ChromEvolOptions (e.g. {10, 3, 1} and {0, 2, 5})
  void optimize(const std::vector<unsigned>& numOfPoints, const std::vector<unsigned>& numOfIterations, double tol,
                double lower = 1e-10, double upper = 100.0) {
    if (paramNames_.empty()) paramNames_ = models_.at(0)->getParameterNames();
    for (size_t i = 0; i < numOfIterations.size(); ++i) {
      sortPoints();
      for (size_t r = 0; r < order_.size(); ++r) active_[order_[r]] = r < numOfPoints[i] ? 1 : 0;   // clearVectorOfLikelihoods (:156-161)
      for (unsigned it = 0; it < numOfIterations[i]; ++it) {
        const Vdouble before = batch_.getValues();
        for (const std::string& name : paramNames_) brent_.search(name, lower, upper, tol, active_);
        double change = 0;
        for (size_t k = 0; k < active_.size(); ++k) if (active_[k]) change = std::max(change, std::fabs(before[k] - batch_.getValue(k)));
        if (change < tol) break;
      }
    }
    sortPoints();
  }
  // points from best to worst (getVectorOfLikelihoods()[0] is the reference's final answer)
  const std::vector<size_t>& getPointOrder() const { return order_; }
  double getValue(size_t point) const { return batch_.getValue(point); }
  double getBestValue() const { return batch_.getValue(order_[0]); }
  ChromosomeSubstitutionModel* getModel(size_t point) const { return models_.at(point); }
  ChromosomeSubstitutionModel* getBestModel() const { return models_.at(order_[0]); }
  unsigned getNumberOfBatchEvaluations() const { return brent_.getNumberOfBatchEvaluations(); }
  LikelihoodPointBatch& getLikelihoods() { return batch_; }

 private:
  struct Adaptor {
    explicit Adaptor(ChromosomeNumberOptimizer* o) : o_(o) {}
    size_t size() const { return o_->models_.size(); }
    double parameter(size_t k, const std::string& name) const { return o_->models_[k]->getParameterValue(name); }
    void setParameter(size_t k, const std::string& name, double v) { o_->models_[k]->setParameterValue(name, v); o_->batch_.markModelChanged(k); }
    void evaluate() { o_->batch_.evaluate(); }
    double value(size_t k) const { return o_->batch_.getValue(k); }
    ChromosomeNumberOptimizer* o_;
  };
  void sortPoints() {
    std::stable_sort(order_.begin(), order_.end(), [&](size_t a, size_t b) {
      if (active_[a] != active_[b]) return active_[a] > active_[b];   // dropped points stay behind the kept ones
      return batch_.getValue(a) < batch_.getValue(b);
    });
  }
  std::vector<ChromosomeSubstitutionModel*> models_;   // not owned
  LikelihoodPointBatch batch_;
  Adaptor adaptor_;
  BatchedBrent<Adaptor> brent_;
  std::vector<char> active_;
  std::vector<size_t> order_;
  std::vector<std::string> paramNames_;
};

// ---- Newton-Raphson on the branch lengths --------------------------------------------------------------------------
// Likelihood/PseudoNewtonOptimizer.cpp:100-193, the optimiser OptimizationTools::optimizeNumericalParameters puts on the branch
// lengths (OptimizationTools.cpp:187-188): every parameter moves by d1 / d2 at once (the other way when d2 < 0, not at all when
// d2 = 0 or the ratio is NaN), and the whole step is halved -- the Felsenstein-Churchill correction -- up to maxCorrection_ times
// while the function is worse than before; stop when |f - f_previous| < tolerance (FunctionStopCondition).  This is the variant
// with disableCG() (PseudoNewtonOptimizer.h:125): no conjugate-gradient detour at the fourth correction.
// One step costs the reference 2 B per-branch derivative passes + the probes; here it is ONE device evaluation with
// BPPGPU_EVAL_D1 | D2 (all B first and second derivatives) + one value evaluation per probe.
class PseudoNewtonOptimizer {
 public:
  explicit PseudoNewtonOptimizer(AbstractHomogeneousTreeLikelihood* function)
      : f_(function), tolerance_(1e-6), maxCorrection_(10), nbEvalMax_(1000000), nbEval_(0), currentValue_(0), previousValue_(0) {}
  void setMaximumNumberOfCorrections(unsigned mx) { maxCorrection_ = mx; }
  void setMaximumNumberOfEvaluations(unsigned n) { nbEvalMax_ = n; }
  void setTolerance(double t) { tolerance_ = t; }
  void disableCG() {}
  void init(const ParameterList& params) {
    params_ = params;
    f_->setParametersValues(params_);
    currentValue_ = f_->getValue();
    previousValue_ = currentValue_;
    nbEval_ = 0;
  }
  double step() {
    const size_t n = params_.size();
    std::vector<double> movements(n);
    ParameterList newPoint = params_;
    for (size_t i = 0; i < n; ++i) {
      const double d1 = f_->getFirstOrderDerivative(params_[i].name), d2 = f_->getSecondOrderDerivative(params_[i].name);
      if (d2 == 0) movements[i] = 0;
      else if (d2 < 0) movements[i] = -d1 / d2;   // "Moving in the other direction" (:121-127)
      else movements[i] = d1 / d2;
      if (std::isnan(movements[i])) movements[i] = 0;
      newPoint[i].value = params_[i].value - movements[i];
    }
    double newValue = probe(newPoint, &movements);
    unsigned count = 0;
    while (count < maxCorrection_ && (newValue > currentValue_ + tolerance_ || std::isnan(newValue))) {
      for (size_t i = 0; i < n; ++i) {
        movements[i] /= 2;
        newPoint[i].value = params_[i].value - movements[i];
      }
      newValue = probe(newPoint, nullptr);
      ++count;
    }
    if (newValue > currentValue_ + tolerance_) {   // "Value could not be ameliorated!"
      f_->setParametersValues(params_);
      newValue = currentValue_;
      previousValue_ = currentValue_;
    } else {
      previousValue_ = currentValue_;
      params_ = newPoint;
      currentValue_ = newValue;
    }
    return newValue;
  }
  unsigned optimize() {
    for (;;) {
      step();
      if (std::fabs(currentValue_ - previousValue_) < tolerance_ || nbEval_ >= nbEvalMax_) break;
    }
    return nbEval_;
  }
  double getFunctionValue() const { return currentValue_; }
  unsigned getNumberOfEvaluations() const { return nbEval_; }
  const ParameterList& getParameters() const { return params_; }

 private:
  // f(newPoint) with the function's own constraints applied (AUTO policy): the point actually reached is read back, and the
  // movement corrected to it like :139-141
  double probe(ParameterList& point, std::vector<double>* movements) {
    f_->setParametersValues(point);
    ++nbEval_;
    for (size_t i = 0; i < point.size(); ++i) {
      point[i].value = f_->getParameterValue(point[i].name);
      if (movements) (*movements)[i] = params_[i].value - point[i].value;
    }
    return f_->getValue();
  }
  AbstractHomogeneousTreeLikelihood* f_;
  ParameterList params_;
  double tolerance_;
  unsigned maxCorrection_, nbEvalMax_, nbEval_;
  double currentValue_, previousValue_;
};

}  // namespace bppshim
