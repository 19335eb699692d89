// boss_b200_fit.h -- the symbolic layer the EM driver needs, and the EM driver itself
// (`boss -T` / baumWelchFit cannot complete without an M-step, and the
// reference's M-step is GSL's BFGS, a system package that is not in this image).
//
// What is mirrored, with the reference's names and JSON formats:
//   WeightExpr / WeightAlgebra   src/weight.h:54-114      expression trees over named parameters
//                                                         (JSON grammar of weight.cpp:547-590)
//   Params / ParamDefs           src/params.h:14-37
//   Constraints                  src/constraints.h:14-28  prob / rate / norm, defaultParams (constraints.cpp:65-75)
//   Machine (basic form)         src/machine.h:93-103     states, transitions, "defs", "cons" (machine.cpp:347-...)
//   EvaluatedMachine(Machine, Params)   src/eval.cpp:42-70
//   MachineCounts::paramCounts   src/counts.cpp:89-106
//   MachineObjective             src/counts.cpp:117-295   same objective and the same re-parameterisation
//                                                         (stick-breaking z = exp(-x^2) for norm groups,
//                                                         p = exp(-x^2) for prob, r = x^2 for rate)
//   MachineFitter::fit           src/fitter.cpp:23-47     same loop and stopping rule
//
// What is NOT mirrored: machine algebra (compose, concat, union, ... machine.cpp:352-430) -- a JSON
// machine using those operators is rejected; build the composite with the reference and pass the
// result.  Derivatives are exact forward-mode (dual numbers) instead of symbolic; the minimiser is a
// plain BFGS with a backtracking line search and the reference's limits (gradient norm 1e-3, 100
// iterations).  The reference's own fit goldens pin the M-step to 4 significant digits only
// (Makefile:502-513), so beyond that M-step parity is unpinned by construction; the E-step is not
// affected.
#ifndef MB_HOST_BOSS_B200_FIT_H
#define MB_HOST_BOSS_B200_FIT_H

#include <algorithm>
#include <cmath>
#include <functional>
#include <set>

#include "boss_b200.h"

namespace MachineBoss {

// ---- weight expressions ----
struct WeightExprNode;
typedef std::shared_ptr<const WeightExprNode> WeightExpr;

struct WeightExprNode {
  enum Op { Const, Param, Mul, Div, Add, Sub, Log, Exp } op = Const;
  double value = 0;
  string name;
  WeightExpr l, r;
};

typedef map<string, WeightExpr> ParamDefs;

struct Dual {   // value + gradient with respect to the active variables
  double v = 0;
  vector<double> d;
  Dual() {}
  Dual (double v_, size_t n) : v (v_), d (n, 0.) {}
};

namespace WeightAlgebra {
  inline WeightExpr doubleConstant (double x) { auto n = std::make_shared<WeightExprNode>(); n->op = WeightExprNode::Const; n->value = x; return n; }
  inline WeightExpr one() { return doubleConstant (1.); }
  inline WeightExpr zero() { return doubleConstant (0.); }
  inline WeightExpr param (const string& p) { auto n = std::make_shared<WeightExprNode>(); n->op = WeightExprNode::Param; n->name = p; return n; }
  inline WeightExpr binary (WeightExprNode::Op op, const WeightExpr& l, const WeightExpr& r) { auto n = std::make_shared<WeightExprNode>(); n->op = op; n->l = l; n->r = r; return n; }
  inline WeightExpr multiply (const WeightExpr& l, const WeightExpr& r) { return binary (WeightExprNode::Mul, l, r); }
  inline WeightExpr divide (const WeightExpr& l, const WeightExpr& r) { return binary (WeightExprNode::Div, l, r); }
  inline WeightExpr add (const WeightExpr& l, const WeightExpr& r) { return binary (WeightExprNode::Add, l, r); }
  inline WeightExpr subtract (const WeightExpr& l, const WeightExpr& r) { return binary (WeightExprNode::Sub, l, r); }
  inline WeightExpr logOf (const WeightExpr& x) { return binary (WeightExprNode::Log, x, nullptr); }
  inline WeightExpr expOf (const WeightExpr& x) { return binary (WeightExprNode::Exp, x, nullptr); }
  inline WeightExpr negate (const WeightExpr& x) { return subtract (one(), x); }          // "not": 1 - x
  inline WeightExpr minus (const WeightExpr& x) { return subtract (zero(), x); }
  inline WeightExpr geometricSum (const WeightExpr& x) { return divide (one(), negate (x)); }   // 1 / (1 - x)

  // JSON grammar of weight.cpp:547-590 ("expr" strings need the reference's PEG parser: rejected)
  inline WeightExpr fromJson (const Json& w) {
    switch (w.type) {
      case Json::Null: return one();
      case Json::Bool: return w.b ? one() : zero();
      case Json::Number: return doubleConstant (w.num);
      case Json::String: return param (w.str);
      case Json::Array: throw runtime_error ("Unexpected type in WeightExpr: array");
      case Json::Object: break;
    }
    if (w.obj.empty()) throw runtime_error ("No opcode in WeightExpr");
    const string& opcode = w.obj.begin()->first;
    const Json& args = w.obj.begin()->second;
    if (opcode == "log") return logOf (fromJson (args));
    if (opcode == "exp") return expOf (fromJson (args));
    if (opcode == "not") return negate (fromJson (args));
    if (opcode == "geomsum") return geometricSum (fromJson (args));
    if (opcode == "*") return multiply (fromJson (args.at (0)), fromJson (args.at (1)));
    if (opcode == "/") return divide (fromJson (args.at (0)), fromJson (args.at (1)));
    if (opcode == "+") return add (fromJson (args.at (0)), fromJson (args.at (1)));
    if (opcode == "-") return subtract (fromJson (args.at (0)), fromJson (args.at (1)));
    throw runtime_error ("Unknown or unsupported opcode " + opcode + " in WeightExpr JSON");
  }

  inline void params (const WeightExpr& w, const ParamDefs& defs, std::set<string>& out, std::set<string>& visiting) {
    if (!w) return;
    if (w->op == WeightExprNode::Param) {
      if (defs.count (w->name)) {
        if (visiting.count (w->name)) throw runtime_error ("Cyclic parameter definition: " + w->name);
        visiting.insert (w->name);
        params (defs.at (w->name), defs, out, visiting);
        visiting.erase (w->name);
      } else out.insert (w->name);
      return;
    }
    params (w->l, defs, out, visiting);
    params (w->r, defs, out, visiting);
  }
  inline std::set<string> params (const WeightExpr& w, const ParamDefs& defs) { std::set<string> out, vis; params (w, defs, out, vis); return out; }

  // value and exact gradient; `vars` maps the active variable names to their index and current value
  struct Env {
    const ParamDefs* defs;
    const map<string, std::pair<size_t, double> >* vars;
    size_t n;
    map<string, Dual> memo;
    int depth = 0;
  };
  inline Dual evalDual (const WeightExpr& w, Env& env) {
    typedef WeightExprNode N;
    if (!w) return Dual (1., env.n);
    switch (w->op) {
      case N::Const: return Dual (w->value, env.n);
      case N::Param: {
        if (env.vars && env.vars->count (w->name)) {
          const auto& iv = env.vars->at (w->name);
          Dual r (iv.second, env.n);
          r.d[iv.first] = 1.;
          return r;
        }
        auto m = env.memo.find (w->name);
        if (m != env.memo.end()) return m->second;
        if (!env.defs || !env.defs->count (w->name)) throw runtime_error ("Parameter " + w->name + " not defined");
        if (++env.depth > 1000) throw runtime_error ("Cyclic parameter definition: " + w->name);
        const Dual r = evalDual (env.defs->at (w->name), env);
        --env.depth;
        env.memo[w->name] = r;
        return r;
      }
      default: break;
    }
    const Dual a = evalDual (w->l, env);
    Dual r (0., env.n);
    if (w->op == N::Log) { r.v = std::log (a.v); for (size_t k = 0; k < env.n; ++k) r.d[k] = a.d[k] / a.v; return r; }
    if (w->op == N::Exp) { r.v = std::exp (a.v); for (size_t k = 0; k < env.n; ++k) r.d[k] = a.d[k] * r.v; return r; }
    const Dual b = evalDual (w->r, env);
    switch (w->op) {
      case N::Mul: r.v = a.v * b.v; for (size_t k = 0; k < env.n; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k]; break;
      case N::Div: r.v = a.v / b.v; for (size_t k = 0; k < env.n; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) / b.v; break;
      case N::Add: r.v = a.v + b.v; for (size_t k = 0; k < env.n; ++k) r.d[k] = a.d[k] + b.d[k]; break;
      case N::Sub: r.v = a.v - b.v; for (size_t k = 0; k < env.n; ++k) r.d[k] = a.d[k] - b.d[k]; break;
      default: break;
    }
    return r;
  }
  inline double eval (const WeightExpr& w, const ParamDefs& defs) {
    Env env { &defs, nullptr, 0 };
    return evalDual (w, env).v;
  }
  inline double asDouble (const WeightExpr& w) {
    if (!w || w->op != WeightExprNode::Const) throw runtime_error ("Parameter value is not a number");
    return w->value;
  }
}  // namespace WeightAlgebra

// ---- Params (src/params.h) ----
struct Params {
  ParamDefs defs;
  void readJson (const Json& pj) { for (const auto& kv: pj.obj) defs[kv.first] = WeightAlgebra::fromJson (kv.second); }
  void writeJson (ostream& out) const {   // numeric parameters, name order (params.cpp)
    out << "{";
    size_t n = 0;
    for (const auto& kv: defs) {
      out << (n++ ? "," : "") << "\"" << Json::escape (kv.first) << "\":";
      if (kv.second && kv.second->op == WeightExprNode::Const) out << kv.second->value;
      else out << WeightAlgebra::eval (kv.second, defs);
    }
    out << "}";
  }
  Params combine (const Params& p, bool overwriteOwnDefs = false) const {   // params.cpp: like JavaScript's extend()
    Params q (*this);
    for (const auto& kv: p.defs) {
      if (q.defs.count (kv.first) && !overwriteOwnDefs) {
        const WeightExpr &a = q.defs.at (kv.first), &b = kv.second;
        const bool same = a && b && a->op == WeightExprNode::Const && b->op == WeightExprNode::Const && a->value == b->value;
        if (!same) throw runtime_error ("Inconsistent parameter definitions for " + kv.first);
      }
      q.defs[kv.first] = kv.second;
    }
    return q;
  }
  static Params fromFile (const string& filename) {
    std::ifstream f (filename);
    if (!f) throw runtime_error ("File not found: " + filename);
    std::stringstream ss; ss << f.rdbuf();
    Params p; p.readJson (Json::parse (ss.str()));
    return p;
  }
};
typedef Params ParamAssign;
typedef Params ParamFuncs;

// ---- Constraints (src/constraints.h) ----
struct Constraints {
  vector<string> prob, rate;
  vector<vector<string> > norm;
  bool empty() const { return prob.empty() && rate.empty() && norm.empty(); }
  void readJson (const Json& pj) {
    if (pj.has ("norm")) for (const auto& n: pj.at ("norm").arr) { vector<string> c; for (const auto& p: n.arr) c.push_back (p.asString()); norm.push_back (c); }
    if (pj.has ("prob")) for (const auto& p: pj.at ("prob").arr) prob.push_back (p.asString());
    if (pj.has ("rate")) for (const auto& r: pj.at ("rate").arr) rate.push_back (r.asString());
  }
  Params defaultParams() const {   // constraints.cpp:65-75
    Params params;
    for (auto& c: norm) for (auto& cp: c) params.defs[cp] = WeightAlgebra::doubleConstant (1. / (double) c.size());
    for (auto& pp: prob) params.defs[pp] = WeightAlgebra::doubleConstant (.5);
    for (auto& rp: rate) params.defs[rp] = WeightAlgebra::doubleConstant (1.);
    return params;
  }
  Constraints combine (const Constraints& c) const {
    Constraints r (*this);
    r.prob.insert (r.prob.end(), c.prob.begin(), c.prob.end());
    r.rate.insert (r.rate.end(), c.rate.begin(), c.rate.end());
    r.norm.insert (r.norm.end(), c.norm.begin(), c.norm.end());
    return r;
  }
  static Constraints fromFile (const string& filename) {
    std::ifstream f (filename);
    if (!f) throw runtime_error ("File not found: " + filename);
    std::stringstream ss; ss << f.rdbuf();
    Constraints c; c.readJson (Json::parse (ss.str()));
    return c;
  }
};

// ---- Machine, basic form (src/machine.h:93-103, machine.cpp readJson's final branch) ----
struct SymbolicTransition {
  InputSymbol in;
  OutputSymbol out;
  StateIndex dest = 0;
  WeightExpr weight;
};

struct MachineState {
  Json name;
  vector<SymbolicTransition> trans;
};

struct Machine {
  ParamFuncs funcs;
  Constraints cons;
  vector<MachineState> state;
  StateIndex nStates() const { return state.size(); }
  void readJson (const Json& pj) {
    static const char* algebra[] = { "compose", "compose-sum", "compose-unsort", "concat", "intersect", "intersect-sum", "intersect-unsort",
                                     "union", "loop", "opt", "star", "plus", "eliminate", "merge", "reverse", "revcomp", "transpose" };
    for (const char* op: algebra)
      if (pj.has (op)) throw runtime_error (string ("machine JSON uses the \"") + op + "\" operator: machine algebra is not part of the B200 path; build the machine with the reference and pass the result");
    if (pj.has ("defs")) funcs.readJson (pj.at ("defs"));
    if (pj.has ("cons")) cons.readJson (pj.at ("cons"));
    const Json& jstate = pj.at ("state");
    map<string, StateIndex> id2n;
    std::set<string> dupIds;
    for (const Json& js: jstate.arr) {
      MachineState ms;
      if (js.has ("n") && (StateIndex) js.at ("n").asInt() != state.size()) throw runtime_error ("StateIndex n out of sequence");
      if (js.has ("id")) {
        const string idStr = js.at ("id").dump();
        if (id2n.count (idStr)) dupIds.insert (idStr); else id2n[idStr] = state.size();
        ms.name = js.at ("id");
      }
      state.push_back (ms);
    }
    size_t n = 0;
    for (const Json& js: jstate.arr) {
      MachineState& ms = state[n++];
      if (!js.has ("trans")) continue;
      for (const Json& jt: js.at ("trans").arr) {
        SymbolicTransition t;
        const Json& dest = jt.at ("to");
        if (dest.type == Json::Number) t.dest = (StateIndex) dest.asInt();
        else {
          const string dstr = dest.dump();
          if (!id2n.count (dstr)) throw runtime_error ("No such state in \"to\": " + dstr);
          if (dupIds.count (dstr)) throw runtime_error ("Ambiguous destination state ID in \"to\": " + dstr);
          t.dest = id2n.at (dstr);
        }
        if (jt.has ("in")) t.in = jt.at ("in").asString();
        if (jt.has ("out")) t.out = jt.at ("out").asString();
        if (jt.has ("expr")) throw runtime_error ("\"expr\" weight strings need the reference's parser; use the JSON weight form");
        t.weight = jt.has ("weight") ? WeightAlgebra::fromJson (jt.at ("weight")) : WeightAlgebra::one();
        if (t.dest >= state.size()) throw runtime_error ("State " + std::to_string (t.dest) + " does not exist");
        ms.trans.push_back (t);
      }
    }
  }
  static Machine fromFile (const string& filename) {
    std::ifstream f (filename);
    if (!f) throw runtime_error ("File not found: " + filename);
    std::stringstream ss; ss << f.rdbuf();
    Machine m; m.readJson (Json::parse (ss.str()));
    return m;
  }
  vector<InputSymbol> inputAlphabet() const {   // alphabetically sorted (machine.cpp:175-182)
    std::set<string> a;
    for (const auto& ms: state) for (const auto& t: ms.trans) if (!t.in.empty()) a.insert (t.in);
    return vector<string> (a.begin(), a.end());
  }
  vector<OutputSymbol> outputAlphabet() const {
    std::set<string> a;
    for (const auto& ms: state) for (const auto& t: ms.trans) if (!t.out.empty()) a.insert (t.out);
    return vector<string> (a.begin(), a.end());
  }
  Params getParamDefs (bool assignDefaultValuesToMissingParams = false) const {   // machine.cpp:2022-2027
    Params p = funcs;
    if (assignDefaultValuesToMissingParams) p = cons.defaultParams().combine (p, true);
    return p;
  }
};

// EvaluatedMachine (Machine, Params): log (WeightAlgebra::eval (weight)) per transition (eval.cpp:42-70)
inline EvaluatedMachine evaluate (const Machine& machine, const Params& params) {
  vector<MachineTransition> flat;
  vector<Json> names;
  for (StateIndex s = 0; s < machine.nStates(); ++s) {
    names.push_back (machine.state[s].name);
    for (const auto& t: machine.state[s].trans) {
      MachineTransition f;
      f.src = s; f.dest = t.dest; f.in = t.in; f.out = t.out;
      f.logWeight = std::log (WeightAlgebra::eval (t.weight, params.defs));
      flat.push_back (f);
    }
  }
  return EvaluatedMachine (machine.nStates(), machine.inputAlphabet(), machine.outputAlphabet(), flat, names);
}

inline vector<double> logWeights (const Machine& machine, const Params& params) {
  vector<double> lw;
  for (const auto& ms: machine.state) for (const auto& t: ms.trans) lw.push_back (std::log (WeightAlgebra::eval (t.weight, params.defs)));
  return lw;
}

// MachineCounts::paramCounts (counts.cpp:89-106): expectation of d(logLike)/d(log param)
inline map<string, double> paramCounts (const MachineCounts& counts, const Machine& machine, const ParamAssign& prob) {
  map<string, double> paramCount;
  for (StateIndex s = 0; s < machine.nStates(); ++s) {
    size_t ti = 0;
    for (const auto& trans: machine.state[s].trans) {
      const double c = counts.count[s][ti++];
      const std::set<string> ps = WeightAlgebra::params (trans.weight, ParamDefs());
      if (ps.empty()) continue;
      map<string, std::pair<size_t, double> > vars;
      size_t k = 0;
      for (const auto& p: ps) { vars[p] = std::make_pair (k++, WeightAlgebra::asDouble (prob.defs.at (p))); }
      WeightAlgebra::Env env { &prob.defs, &vars, ps.size() };
      const Dual w = WeightAlgebra::evalDual (trans.weight, env);
      for (const auto& p: ps) paramCount[p] += c * w.d[vars[p].first] * vars[p].second / w.v;
    }
  }
  return paramCount;
}

// ---- M-step (the role of MachineObjective, counts.cpp:108-295) ----
// Maximise  Q(theta) = sum_t count_t * log w_t(theta)  over the constrained parameters theta.  Every degree of
// freedom gets an unconstrained coordinate u, and each constrained parameter is an expression of coordinates:
//   "prob"  p = exp(-u^2)          "rate"  r = u^2
//   "norm"  group p_0 .. p_{n-1}: a unit stick is broken n-1 times, break k keeping the fraction
//           keep_k = exp(-u_k^2) of what is left:   p_k = (1 - keep_k) prod_{j<k} keep_j,   p_{n-1} = prod_j keep_j
// This is the reference's map (it has to be: its fit goldens pin the optimum's basin to 4 digits, and BFGS from
// the same start in the same coordinates is what reproduces them); everything else here -- the objective kept
// as a list of (count, weight) terms and differentiated with dual numbers, the coordinate bookkeeping, the
// minimiser -- is this file's own.
struct MachineObjective {
  struct Term { double count; WeightExpr weight; };
  struct Group { vector<string> member; size_t firstCoord; };      // a norm group and its n-1 coordinates
  struct Single { string name; size_t coord; bool isRate; };       // a prob or rate parameter and its coordinate

  vector<Term> terms;              // transitions with a non-zero expected count
  vector<Group> groups;
  vector<Single> singles;
  vector<string> coordName;        // coordinate k is the variable coordName[k]
  ParamDefs bound;                 // constants and functions, plus every constrained parameter in coordinates

  MachineObjective (const Machine& machine, const MachineCounts& counts, const Constraints& cons, const Params& constants) {
    using namespace WeightAlgebra;
    std::set<string> taken;        // names already meaning something: a coordinate must not shadow them
    for (StateIndex s = 0; s < machine.nStates(); ++s)
      for (size_t t = 0; t < machine.state[s].trans.size(); ++t) {
        const WeightExpr& w = machine.state[s].trans[t].weight;
        for (const auto& nm: params (w, ParamDefs())) taken.insert (nm);
        if (counts.count[s][t] != 0.) terms.push_back (Term { counts.count[s][t], w });
      }
    bound = machine.funcs.combine (constants).defs;
    for (const auto& kv: bound) taken.insert (kv.first);
    size_t serial = 0;
    auto newCoord = [&] () {
      string nm;
      do nm = "$x" + std::to_string (++serial); while (taken.count (nm));
      coordName.push_back (nm);
      return coordName.size() - 1;
    };
    auto keepOf = [&] (size_t coord) {      // exp(-u^2)
      const WeightExpr u = param (coordName[coord]);
      return expOf (minus (multiply (u, u)));
    };
    const Constraints all = machine.cons.combine (cons);
    for (const auto& members: all.norm) {
      if (members.empty()) continue;
      Group g { members, coordName.size() };
      for (size_t k = 0; k + 1 < members.size(); ++k) newCoord();
      WeightExpr left = one();               // what remains of the stick before break k
      for (size_t k = 0; k < members.size(); ++k) {
        if (k + 1 == members.size()) { bound[members[k]] = left; break; }
        const WeightExpr keep = keepOf (g.firstCoord + k);
        bound[members[k]] = multiply (left, negate (keep));
        left = multiply (left, keep);
      }
      groups.push_back (g);
    }
    for (const auto& nm: all.prob) { const size_t c = newCoord(); bound[nm] = keepOf (c); singles.push_back (Single { nm, c, false }); }
    for (const auto& nm: all.rate) { const size_t c = newCoord(); const WeightExpr u = param (coordName[c]); bound[nm] = multiply (u, u); singles.push_back (Single { nm, c, true }); }
  }

  // -Q and its gradient at the coordinates u (exact forward-mode derivatives)
  double eval (const vector<double>& u, vector<double>* grad) const {
    map<string, std::pair<size_t, double> > vars;
    for (size_t k = 0; k < coordName.size(); ++k) vars[coordName[k]] = std::make_pair (k, u[k]);
    WeightAlgebra::Env env { &bound, &vars, u.size() };      // (env.memo shares the parameter expressions between terms)
    double f = 0;
    if (grad) grad->assign (u.size(), 0.);
    for (const Term& t: terms) {
      const Dual w = WeightAlgebra::evalDual (t.weight, env);
      f -= t.count * std::log (w.v);
      if (grad) for (size_t k = 0; k < u.size(); ++k) (*grad)[k] -= t.count * w.d[k] / w.v;
    }
    return f;
  }

  // coordinates of a parameter assignment: the inverse of the map above
  vector<double> coordinatesOf (const Params& at) const {
    vector<double> u (coordName.size(), 0.);
    auto value = [&] (const string& nm) { return WeightAlgebra::asDouble (at.defs.at (nm)); };
    for (const Group& g: groups) {
      double left = 1;
      for (size_t k = 0; k + 1 < g.member.size(); ++k) {
        const double keep = 1 - value (g.member[k]) / left;      // fraction of the remaining stick that survives break k
        u[g.firstCoord + k] = std::sqrt (-std::log (keep));
        left -= value (g.member[k]);
      }
    }
    for (const Single& sp: singles) u[sp.coord] = sp.isRate ? std::sqrt (value (sp.name)) : std::sqrt (-std::log (value (sp.name)));
    return u;
  }

  Params optimize (const Params& seed) const {
    vector<double> u = coordinatesOf (seed);
    if (!u.empty()) bfgs (u);
    Params fitted = seed;
    map<string, std::pair<size_t, double> > vars;
    for (size_t k = 0; k < u.size(); ++k) vars[coordName[k]] = std::make_pair (k, u[k]);
    WeightAlgebra::Env env { &bound, &vars, u.size() };
    auto assign = [&] (const string& nm) { fitted.defs[nm] = WeightAlgebra::doubleConstant (WeightAlgebra::evalDual (bound.at (nm), env).v); };
    for (const Group& g: groups) for (const auto& nm: g.member) assign (nm);
    for (const Single& sp: singles) assign (sp.name);
    return fitted;
  }

private:
  // BFGS on the inverse Hessian with an Armijo backtracking line search; the reference's limits:
  // stop when |gradient| < 1e-3 (EpsilonAbsolute) or after 100 iterations (MaxIterations).
  void bfgs (vector<double>& x) const {
    const size_t n = x.size();
    vector<double> g (n), gNew (n), d (n), xNew (n), s (n), yv (n);
    vector<vector<double> > H (n, vector<double> (n, 0.));
    for (size_t k = 0; k < n; ++k) H[k][k] = 1.;
    double f = eval (x, &g);
    auto norm = [] (const vector<double>& v) { double t = 0; for (double e: v) t += e * e; return std::sqrt (t); };
    for (int iter = 0; iter < 100; ++iter) {
      if (!std::isfinite (f)) break;
      for (size_t i = 0; i < n; ++i) { d[i] = 0; for (size_t j = 0; j < n; ++j) d[i] -= H[i][j] * g[j]; }
      double slope = 0;
      for (size_t i = 0; i < n; ++i) slope += d[i] * g[i];
      if (!(slope < 0)) { for (size_t i = 0; i < n; ++i) { d[i] = -g[i]; for (size_t j = 0; j < n; ++j) H[i][j] = (i == j); } slope = -norm (g) * norm (g); }
      // first step like gsl's bfgs2 (StepSize 0.1 along the unit direction), then unit steps
      double alpha = iter == 0 ? std::min (1., 0.1 / std::max (norm (d), 1e-300)) : 1.;
      double fNew = f;
      bool ok = false;
      for (int ls = 0; ls < 60; ++ls) {
        for (size_t i = 0; i < n; ++i) xNew[i] = x[i] + alpha * d[i];
        fNew = eval (xNew, &gNew);
        if (std::isfinite (fNew) && fNew <= f + 1e-4 * alpha * slope) { ok = true; break; }
        alpha *= 0.5;
      }
      if (!ok) break;
      double ys = 0;
      for (size_t i = 0; i < n; ++i) { s[i] = xNew[i] - x[i]; yv[i] = gNew[i] - g[i]; ys += s[i] * yv[i]; }
      if (ys > 1e-300) {   // H <- (I - s y'/ys) H (I - y s'/ys) + s s'/ys
        vector<double> Hy (n, 0.);
        for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) Hy[i] += H[i][j] * yv[j];
        double yHy = 0;
        for (size_t i = 0; i < n; ++i) yHy += yv[i] * Hy[i];
        for (size_t i = 0; i < n; ++i)
          for (size_t j = 0; j < n; ++j)
            H[i][j] += (1 + yHy / ys) * s[i] * s[j] / ys - (Hy[i] * s[j] + s[i] * Hy[j]) / ys;
      }
      x = xNew; g = gNew; f = fNew;
      if (norm (g) < 1e-3) break;
    }
  }
};

// ---- EM driver (src/fitter.h:11-20, fitter.cpp:23-47) ----
struct MachineFitter {
  Machine machine;
  Constraints constraints;
  Params seed, constants;
  int iterations = 0;             // EM iterations run by the last fit()
  double logLike = 0;             // log-likelihood of the last E-step
  std::function<void (int, double, const Params&)> onIteration;   // the reference logs these at -v2 / -v4

  Constraints allConstraints() const { return machine.cons.combine (constraints); }

  // fitter.h:18-20.  As in the reference, the envelopes handed in reach MachineCounts::add (counts.cpp:57-64) and
  // stop there: the matrices build their own from each SeqPair (dpmatrix.defs.h:15-27), i.e. the path envelope
  // for a pair that carries an alignment and the full matrix otherwise -- which is what DeviceBatch sets up.
  Params fit (const SeqPairList& trainingSet, size_t width) { return fit (trainingSet, envelopes (trainingSet, width)); }
  Params fit (const SeqPairList& trainingSet, const list<Envelope>& envs) {
    if (envs.size() != trainingSet.seqPairs.size()) throw runtime_error ("Envelope/training set mismatch");      // fitter.cpp:24
    return fit (trainingSet);
  }
  Params fit (const SeqPairList& trainingSet) {
    const int MaxEMIterations = 1000;
    const double MinEMImprovement = .001;
    Params params = seed;
    double prev = 0;
    // one device copy of the machine structure and of the batch for the whole fit: only the weights change
    EvaluatedMachine eval = evaluate (machine, machine.funcs.combine (constants).combine (params));
    ListBatch batch (eval, pairPointers (trainingSet));      // every visible GPU: pairs dealt to the devices, counts all-reduced
    vector<double> c (eval.nTransitions ? eval.nTransitions : 1), ll (trainingSet.seqPairs.size() ? trainingSet.seqPairs.size() : 1);
    for (int iter = 0; true; ++iter) {
      const Params allParams = machine.funcs.combine (constants).combine (params);
      if (iter > 0) eval.setLogWeights (logWeights (machine, allParams));
      batch.counts (c.data(), ll.data());
      MachineCounts counts (eval);
      for (StateIndex s = 0; s < eval.nStates(); ++s) for (size_t t = 0; t < counts.count[s].size(); ++t) counts.count[s][t] = c[eval.state[s].transOffset + t];
      counts.loglike = 0;
      for (size_t k = 0; k < trainingSet.seqPairs.size(); ++k) counts.loglike += ll[k];
      iterations = iter + 1;
      logLike = counts.loglike;
      if (onIteration) onIteration (iter + 1, counts.loglike, params);
      if (iter > 0) {
        if (iter == MaxEMIterations) break;
        const double improvement = (counts.loglike - prev) / std::fabs (prev);
        if (improvement < MinEMImprovement) break;
      }
      const MachineObjective objective (machine, counts, constraints, constants);
      params = objective.optimize (params);
      prev = counts.loglike;
    }
    return params;
  }
};

// ---- the free functions of src/api.h:14-34, with the reference's signatures: (Machine, Params, data) ----
// Like api.cpp:31-75 each call evaluates the machine anew; the forms taking an EvaluatedMachine (boss_b200.h) skip
// that.  A call on ONE pair pays a device round trip for a single matrix: lists belong in forwardBackwardCounts
// (Machine, Params, SeqPairList), forwardLogLikes / viterbiLogLikes, or MachineFitter.
// The selection of Machine::downsample (machine.cpp:2036-2082): which transitions of an acyclic, topologically sorted machine
// stay when only the most probable ones are kept.  All labels are dropped (the "null" machine), Forward and Backward are filled
// for the empty sequence pair -- one cell, every transition silent -- and the (cell, transition) posteriors are taken from the
// top of BackwardMatrix::postTransQueue: each is traced back to the start and on to the end through the best transitions, every
// transition on the way is kept, a trace stops where it meets one already kept; until maxProportion of the transitions are kept
// or the next posterior falls below minPostProb.  Returns allowed[state][transIndex].  (What the reference does with the mask
// afterwards -- subgraph, ergodicMachine, eliminateRedundantStates -- is machine algebra, outside this path.)
inline vector<vector<bool>> downsampleTransitions (const Machine& machine, double maxProportionOfTransitionsToKeep, double minPostProbOfSelectedTransitions = 0.) {
  Machine null (machine);
  vector<vector<bool>> transAllowed;
  size_t nTransNull = 0;
  for (auto& ms: null.state) {
    for (auto& mt: ms.trans) { mt.in = mt.out = string(); if (mt.dest <= (StateIndex) (&ms - &null.state[0])) throw runtime_error ("Machine must be acyclic & topologically sorted before downsampling can take place"); }
    transAllowed.push_back (vector<bool> (ms.trans.size(), false));
    nTransNull += ms.trans.size();
  }
  const SeqPair emptySeqPair;
  const EvaluatedMachine eval = evaluate (null, machine.getParamDefs (true));
  const ForwardMatrix fwd (eval, emptySeqPair);
  const BackwardMatrix back (eval, emptySeqPair);
  size_t nTrans = 0;
  StoredMatrix::TraceTerminator stopTrace = [&] (long, long, StateIndex s, size_t ti) {
    if (transAllowed[s][ti]) return true;
    transAllowed[s][ti] = true;
    ++nTrans;
    return false;
  };
  BackwardMatrix::PostTransQueue queue = back.postTransQueue (fwd);
  const size_t nTransTarget = (size_t) ((double) nTransNull * maxProportionOfTransitionsToKeep);
  while (!queue.empty() && (nTrans == 0 || nTrans < nTransTarget)) {
    const BackwardMatrix::PostTrans pt = queue.top();
    if (pt.weight < minPostProbOfSelectedTransitions && nTrans > 0) break;
    queue.pop();
    back.traceFrom (fwd, pt.inPos, pt.outPos, pt.src, pt.transIndex, stopTrace);
  }
  return transAllowed;
}

// The selection of Machine::stochasticDownsample (machine.cpp:2084-2128): paths are drawn through the null machine's Forward matrix
// (randomTransSelector, the caller's mt19937) until maxProportion of the transitions lie on a sampled path or maxPaths are drawn.
inline vector<vector<bool>> stochasticDownsampleTransitions (const Machine& machine, std::mt19937& rng, double maxProportionOfTransitionsToKeep, int maxNumberOfPathsToSample) {
  Machine null (machine);
  vector<vector<bool>> transAllowed;
  size_t nTransNull = 0;
  for (auto& ms: null.state) {
    for (auto& mt: ms.trans) { mt.in = mt.out = string(); if (mt.dest <= (StateIndex) (&ms - &null.state[0])) throw runtime_error ("Machine must be acyclic & topologically sorted before stochastic downsampling can take place"); }
    transAllowed.push_back (vector<bool> (ms.trans.size(), false));
    nTransNull += ms.trans.size();
  }
  const size_t nTransTarget = (size_t) ((double) nTransNull * maxProportionOfTransitionsToKeep);
  const SeqPair emptySeqPair;
  const EvaluatedMachine eval = evaluate (null, machine.getParamDefs (true));
  const ForwardMatrix fwd (eval, emptySeqPair);
  size_t nTrans = 0;
  StoredMatrix::TraceTerminator neverStopTrace = [&] (long, long, StateIndex s, size_t ti) {
    if (!transAllowed[s][ti]) { transAllowed[s][ti] = true; ++nTrans; }
    return false;
  };
  const StoredMatrix::TransSelector selectRandomTrans = ForwardMatrix::randomTransSelector (rng);
  for (int nPath = 0; nPath < maxNumberOfPathsToSample && nTrans < nTransTarget; ++nPath)
    fwd.traceBack (0, 0, (StateIndex) null.nStates() - 1, neverStopTrace, selectRandomTrans);
  return transAllowed;
}

inline Machine loadMachine (const string& filename) { return Machine::fromFile (filename); }                              // api.h:15
inline Machine loadMachineJson (const string& jsonString) { Machine m; m.readJson (Json::parse (jsonString)); return m; }  // api.h:16
inline double forwardLogLike (const Machine& machine, const Params& params, const SeqPair& seqPair) {                      // api.h:21
  const EvaluatedMachine eval = evaluate (machine, params);
  return ForwardMatrix (eval, seqPair).logLike();
}
inline double forwardLogLike (const Machine& machine, const Params& params, const SeqPair& seqPair, const Envelope& env) { // api.h:22
  const EvaluatedMachine eval = evaluate (machine, params);
  return ForwardMatrix (eval, seqPair, env).logLike();
}
inline double viterbiLogLike (const Machine& machine, const Params& params, const SeqPair& seqPair) {                      // api.h:25
  const EvaluatedMachine eval = evaluate (machine, params);
  return ViterbiMatrix (eval, seqPair).logLike();
}
inline MachinePath viterbiAlign (const Machine& machine, const Params& params, const SeqPair& seqPair) {                   // api.h:26
  const EvaluatedMachine eval = evaluate (machine, params);
  return ViterbiMatrix (eval, seqPair).path (machine);
}
inline MachineCounts forwardBackwardCounts (const Machine& machine, const Params& params, const SeqPair& seqPair) {        // api.h:29
  const EvaluatedMachine eval = evaluate (machine, params);
  return MachineCounts (eval, seqPair);
}
inline MachineCounts forwardBackwardCounts (const Machine& machine, const Params& params, const SeqPairList& seqPairList) { // api.h:30
  const EvaluatedMachine eval = evaluate (machine, params);
  return MachineCounts (eval, seqPairList);
}

// api.h:33-34
inline Params baumWelchFit (const Machine& machine, const Constraints& constraints, const SeqPairList& data,
                            const Params& seed = Params(), const Params& constants = Params()) {
  MachineFitter fitter;
  fitter.machine = machine;
  fitter.constraints = constraints;
  fitter.constants = constants;
  fitter.seed = fitter.allConstraints().defaultParams().combine (seed, true);
  return fitter.fit (data);
}

}  // namespace MachineBoss

#endif
