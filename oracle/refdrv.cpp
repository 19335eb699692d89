// oracle/refdrv.cpp -- TEST INFRASTRUCTURE (not product code).
//
// A small driver that links the UNMODIFIED reference sources (compiled where they lie under
// /root/reference by oracle/Makefile, objects in oracle/_ref/) and runs the reference's own
// Forward / Backward / Viterbi / counts classes.  It stands in for target/boss.cpp, which needs
// the real Boost program_options (absent from this image).  It mirrors the CLI blocks at
// target/boss.cpp:789-848:
//    -L  -> RollingOutputForwardMatrix::logLike         (boss.cpp:799)
//    -V/-A -> ViterbiMatrix::logLike + traceBack          (boss.cpp:828-833)
//    -C  -> MachineCounts::add = Forward + Backward + getCounts   (counts.cpp:57-64)
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may execute this binary.
//
// Output is JSON on stdout; doubles are printed with 17 significant digits.
//
// Usage:
//   refdrv --machine SPEC [--machine SPEC ...] [--params FILE] [--no-defaults]
//          ( --emit-machine
//          | --downsample SIZE,PROB   (the machine is toposorted first, as boss --downsample-size does, boss.cpp:487-490: prints
//                                      the toposorted symbolic machine and which transitions Machine::downsample keeps)
//          | [--pairs FILE | --synth N,LI,LO,SEED] --do forward,rolling,viterbi,path,backward,counts,matrices
//            [--threads T] [--quiet-results] )
//   SPEC = preset:NAME | file:PATH | hmmer:PATH | hmmer-global:PATH ; several SPECs are composed
//          left to right with Machine::compose, like `boss a b c` (boss.cpp:269-277).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <thread>
#include <vector>

#include "machine.h"
#include "preset.h"
#include "hmmer.h"
#include "eval.h"
#include "seqpair.h"
#include "forward.h"
#include "backward.h"
#include "viterbi.h"
#include "counts.h"
#include "synth.h"

using namespace MachineBoss;
using namespace std;

static string dstr (double x) {
  if (x == numeric_limits<double>::infinity()) return "\"Infinity\"";
  if (x == -numeric_limits<double>::infinity()) return "\"-Infinity\"";
  if (x != x) return "\"NaN\"";
  char buf[64];
  snprintf (buf, sizeof buf, "%.17g", x);
  return buf;
}

static string jstr (const string& s) { return json (s).dump(); }

static Machine loadSpec (const string& spec) {
  const size_t c = spec.find (':');
  const string kind = c == string::npos ? string ("file") : spec.substr (0, c);
  const string arg = c == string::npos ? spec : spec.substr (c + 1);
  if (kind == "preset") return MachinePresets::makePreset (arg);
  if (kind == "file") return MachineLoader::fromFile (arg);
  if (kind == "generate") return Machine::generator (splitToChars (arg), arg);      // boss.cpp:362-364
  if (kind == "recognize") return Machine::recognizer (splitToChars (arg), arg);    // boss.cpp:384-386
  if (kind == "hmmer" || kind == "hmmer-global") {
    HmmerModel hmm;
    ifstream in (arg);
    if (!in) { cerr << "can't open " << arg << endl; exit (1); }
    hmm.read (in);
    return hmm.machine (kind == "hmmer");
  }
  cerr << "unknown machine spec " << spec << endl;
  exit (1);
}

struct PairResult {
  bool tokenizable = true;
  double forward = 0, rolling = 0, viterbi = 0, backward = 0;
  vector<long long> path;  // global transition ids, start -> end
  vector<long long> sample;  // a path drawn by ForwardMatrix::samplePath
  string matrices;
  string postTrans;          // top of BackwardMatrix::postTransQueue + the trace from its first entry
};

int main (int argc, char** argv) {
  vector<string> specs;
  string paramsFile, pairsFile, doList;
  bool useDefaults = true, emitMachine = false, quiet = false;
  string downsampleSpec, downsamplePathSpec;
  long long synthN = 0, synthLi = 0, synthLo = 0, synthSeed = 0;
  int nThreads = 1;
  long long sampleSeed = 1;
  for (int a = 1; a < argc; ++a) {
    const string f = argv[a];
    auto next = [&] () -> string { if (a + 1 >= argc) { cerr << "missing value for " << f << endl; exit (1); } return argv[++a]; };
    if (f == "--machine") specs.push_back (next());
    else if (f == "--params") paramsFile = next();
    else if (f == "--no-defaults") useDefaults = false;
    else if (f == "--emit-machine") emitMachine = true;
    else if (f == "--downsample") downsampleSpec = next();
    else if (f == "--downsample-path") downsamplePathSpec = next();      // FRACTION,PATHS,SEED: Machine::stochasticDownsample's selection
    else if (f == "--pairs") pairsFile = next();
    else if (f == "--synth") { if (sscanf (next().c_str(), "%lld,%lld,%lld,%lld", &synthN, &synthLi, &synthLo, &synthSeed) != 4) { cerr << "bad --synth" << endl; exit (1); } }
    else if (f == "--do") doList = next();
    else if (f == "--threads") nThreads = atoi (next().c_str());
    else if (f == "--sample-seed") sampleSeed = atoll (next().c_str());
    else if (f == "--quiet-results") quiet = true;
    else { cerr << "unknown flag " << f << endl; exit (1); }
  }
  if (specs.empty()) { cerr << "need --machine" << endl; exit (1); }

  try {
    Machine machine = loadSpec (specs[0]);
    for (size_t n = 1; n < specs.size(); ++n)
      machine = Machine::compose (machine, loadSpec (specs[n]));

    // boss.cpp:789: params = funcs ∪ seed ∪ machine.getParamDefs(-U)
    Params seed;
    if (paramsFile.size())
      seed = JsonLoader<ParamAssign>::fromFile (paramsFile);

    if (downsamplePathSpec.size()) {
      // The selection loop of Machine::stochasticDownsample (machine.cpp:2084-2128): paths sampled through the null machine's
      // Forward matrix with randomTransSelector (mt19937 (SEED)) until the fraction of transitions is covered or PATHS are drawn.
      double maxProportion = 1;
      int maxPaths = 1;
      long long seedValue = 1;
      if (sscanf (downsamplePathSpec.c_str(), "%lf,%d,%lld", &maxProportion, &maxPaths, &seedValue) != 3) { cerr << "bad --downsample-path" << endl; exit (1); }
      Machine sorted = machine.toposort();
      sorted.funcs = sorted.funcs.combine (seed, true);
      Machine null (sorted);
      vguard<vguard<bool> > transAllowed;
      for (auto& ms: null.state) {
        for (auto& mt: ms.trans) mt.in = mt.out = string();
        transAllowed.push_back (vguard<bool> (ms.trans.size()));
      }
      const size_t nTransTarget = null.nTransitions() * maxProportion;
      const SeqPair emptySeqPair;
      const EvaluatedMachine nullEval (null, sorted.getParamDefs (true));
      const ForwardMatrix fwd (nullEval, emptySeqPair);
      size_t nTrans = 0, nPath = 0;
      DPMatrix<IdentityIndexMapper>::TraceTerminator neverStopTrace = [&] (Envelope::InputIndex, Envelope::OutputIndex, StateIndex st, EvaluatedMachineState::TransIndex ti) {
        if (!transAllowed[st][ti]) { transAllowed[st][ti] = true; ++nTrans; }
        return false;
      };
      mt19937 rng ((unsigned) seedValue);
      ForwardMatrix::TransSelector selectRandomTrans = fwd.randomTransSelector (rng);
      for (; nPath < (size_t) maxPaths && nTrans < nTransTarget; ++nPath)
        fwd.traceBack (null, fwd.inLen, fwd.outLen, null.endState(), neverStopTrace, selectRandomTrans);
      cout << "{\"machine\":";
      sorted.writeJson (cout, false, true);
      cout << ",\n \"nTransitions\":" << null.nTransitions() << ",\"kept\":" << nTrans << ",\"paths\":" << nPath << ",\n \"allowed\":[";
      for (size_t st = 0; st < transAllowed.size(); ++st) {
        cout << (st ? "," : "") << "[";
        for (size_t ti = 0; ti < transAllowed[st].size(); ++ti) cout << (ti ? "," : "") << (transAllowed[st][ti] ? 1 : 0);
        cout << "]";
      }
      cout << "]}" << endl;
      return 0;
    }

    if (downsampleSpec.size()) {
      // The selection loop of Machine::downsample (machine.cpp:2036-2082) through the reference's own primitives (the null machine,
      // ForwardMatrix / BackwardMatrix of the empty pair, postTransQueue, traceFrom with the stop terminator), so that the mask it
      // builds can be seen: the function itself only returns the machine after subgraph / ergodicMachine / eliminateRedundantStates.
      double maxProportion = 1, minPostProb = 0;
      if (sscanf (downsampleSpec.c_str(), "%lf,%lf", &maxProportion, &minPostProb) != 2) { cerr << "bad --downsample" << endl; exit (1); }
      Machine sorted = machine.toposort();
      sorted.funcs = sorted.funcs.combine (seed, true);      // the parameter values travel with the machine ("defs")
      Machine null (sorted);
      vguard<vguard<bool> > transAllowed;
      for (auto& ms: null.state) {
        for (auto& mt: ms.trans) mt.in = mt.out = string();
        transAllowed.push_back (vguard<bool> (ms.trans.size()));
      }
      const SeqPair emptySeqPair;
      const EvaluatedMachine nullEval (null, sorted.getParamDefs (true));
      const ForwardMatrix fwd (nullEval, emptySeqPair);
      const BackwardMatrix back (nullEval, emptySeqPair);
      size_t nTrans = 0;
      DPMatrix<IdentityIndexMapper>::TraceTerminator stopTrace = [&] (Envelope::InputIndex, Envelope::OutputIndex, StateIndex st, EvaluatedMachineState::TransIndex ti) {
        if (transAllowed[st][ti]) return true;
        transAllowed[st][ti] = true;
        ++nTrans;
        return false;
      };
      BackwardMatrix::PostTransQueue queue = back.postTransQueue (fwd);
      const size_t nTransTarget = null.nTransitions() * maxProportion;
      while (!queue.empty() && (nTrans == 0 || nTrans < nTransTarget)) {
        const BackwardMatrix::PostTrans pt = queue.top();
        if (pt.weight < minPostProb && nTrans > 0) break;
        queue.pop();
        back.traceFrom (null, fwd, pt.inPos, pt.outPos, pt.src, pt.transIndex, stopTrace);
      }
      const Machine kept = sorted.downsample (maxProportion, minPostProb);      // the reference's own call, for its transition count
      cout << "{\"machine\":";
      sorted.writeJson (cout, false, true);
      cout << ",\n \"nTransitions\":" << null.nTransitions() << ",\"kept\":" << nTrans << ",\"downsampledMachineTransitions\":" << kept.nTransitions() << ",\n \"allowed\":[";
      for (size_t st = 0; st < transAllowed.size(); ++st) {
        cout << (st ? "," : "") << "[";
        for (size_t ti = 0; ti < transAllowed[st].size(); ++ti) cout << (ti ? "," : "") << (transAllowed[st][ti] ? 1 : 0);
        cout << "]";
      }
      cout << "]}" << endl;
      return 0;
    }
    const Params params = machine.getParamDefs (useDefaults).combine (seed, true);   // --params overrides the -U defaults
    const EvaluatedMachine eval (machine, params);

    if (emitMachine) {
      // Flat form, in the reference's enumeration order (eval.cpp:49-69): state s ascending, then
      // position in s's TransList.  Global transition id = transOffset[s] + transIndex.
      cout << "{\"nStates\":" << eval.nStates() << ",\n \"inAlphabet\":[";
      for (size_t t = 1; t < eval.inputTokenizer.tok2sym.size(); ++t) cout << (t > 1 ? "," : "") << jstr (eval.inputTokenizer.tok2sym[t]);
      cout << "],\n \"outAlphabet\":[";
      for (size_t t = 1; t < eval.outputTokenizer.tok2sym.size(); ++t) cout << (t > 1 ? "," : "") << jstr (eval.outputTokenizer.tok2sym[t]);
      cout << "],\n \"stateNames\":[";
      for (StateIndex s = 0; s < eval.nStates(); ++s) cout << (s ? "," : "") << machine.state[s].name;   // raw JSON names, null allowed
      cout << "],\n \"trans\":[";
      size_t n = 0;
      for (StateIndex s = 0; s < machine.nStates(); ++s) {
        size_t ti = 0;
        for (const auto& t: machine.state[s].trans) {
          cout << (n++ ? ",\n  " : "\n  ") << "[" << s << "," << t.dest << ","
               << eval.inputTokenizer.sym2tok.at (t.in) << "," << eval.outputTokenizer.sym2tok.at (t.out) << ","
               << dstr (eval.state[s].logTransWeight[ti]) << "," << ti << "]";
          ++ti;
        }
      }
      cout << "\n ]}" << endl;
      return 0;
    }

    // data
    vector<SeqPair> pairs;
    if (pairsFile.size()) {
      SeqPairList data = JsonLoader<SeqPairList>::fromFile (pairsFile);
      pairs.assign (data.seqPairs.begin(), data.seqPairs.end());
    }
    const int nIn = (int) eval.inputTokenizer.tok2sym.size() - 1, nOut = (int) eval.outputTokenizer.tok2sym.size() - 1;
    for (long long k = 0; k < synthN; ++k) {
      SeqPair sp;
      sp.input.name = "x" + to_string (k);
      sp.output.name = "y" + to_string (k);
      if (nIn) for (long long p = 0; p < synthLi; ++p) sp.input.seq.push_back (eval.inputTokenizer.tok2sym[mb_synth_token (synthSeed, k, 0, p, nIn)]);
      if (nOut) for (long long p = 0; p < synthLo; ++p) sp.output.seq.push_back (eval.outputTokenizer.tok2sym[mb_synth_token (synthSeed, k, 1, p, nOut)]);
      pairs.push_back (sp);
    }

    auto wants = [&] (const char* w) { return (("," + doList + ",").find (string (",") + w + ",")) != string::npos; };
    const bool doForward = wants ("forward"), doRolling = wants ("rolling"), doViterbi = wants ("viterbi"), doPath = wants ("path"),
      doBackward = wants ("backward"), doCounts = wants ("counts"), doMatrices = wants ("matrices"), doSample = wants ("sample"),
      doPostTrans = wants ("posttrans");

    vector<PairResult> res (pairs.size());
    vector<MachineCounts> threadCounts ((size_t) nThreads, MachineCounts (eval));
    vector<double> cellStates ((size_t) nThreads, 0.);

    auto work = [&] (int tid) {
      for (size_t k = (size_t) tid; k < pairs.size(); k += (size_t) nThreads) {
        const SeqPair& sp = pairs[k];
        PairResult& r = res[k];
        r.tokenizable = eval.canTokenize (sp);
        if (!r.tokenizable) continue;
        cellStates[tid] += (double) (sp.input.seq.size() + 1) * (double) (sp.output.seq.size() + 1) * (double) eval.nStates();
        if (doRolling) { const RollingOutputForwardMatrix f (eval, sp); r.rolling = f.logLike(); }
        if (doForward && !doCounts && !doMatrices) { const ForwardMatrix f (eval, sp); r.forward = f.logLike(); }
        if (doViterbi || doPath) {
          const ViterbiMatrix v (eval, sp);
          r.viterbi = v.logLike();
          if (doPath && r.viterbi > -numeric_limits<double>::infinity()) {
            ViterbiMatrix::TraceTerminator collect = [&] (Envelope::InputIndex, Envelope::OutputIndex, StateIndex src, EvaluatedMachineState::TransIndex ti) {
              r.path.push_back ((long long) (eval.state[src].transOffset + ti));
              return false;
            };
            v.traceBack (machine, v.inLen, v.outLen, v.nStates - 1, collect);
            std::reverse (r.path.begin(), r.path.end());
          }
        }
        if (doSample) {
          // ForwardMatrix::samplePath (forward.cpp:17-19) = traceBack with randomTransSelector; pair k draws from mt19937 (sampleSeed + k)
          const ForwardMatrix f (eval, sp);
          r.forward = f.logLike();
          if (r.forward > -numeric_limits<double>::infinity()) {
            mt19937 rng ((unsigned) (sampleSeed + (long long) k));
            ForwardMatrix::TraceTerminator collect = [&] (Envelope::InputIndex, Envelope::OutputIndex, StateIndex src, EvaluatedMachineState::TransIndex ti) {
              r.sample.push_back ((long long) (eval.state[src].transOffset + ti));
              return false;
            };
            f.traceBack (machine, f.inLen, f.outLen, f.nStates - 1, collect, ForwardMatrix::randomTransSelector (rng));
            std::reverse (r.sample.begin(), r.sample.end());
          }
        }
        if (doCounts) {
          // counts.cpp:57-64
          const ForwardMatrix f (eval, sp);
          const BackwardMatrix b (eval, sp);
          b.getCounts (f, threadCounts[tid]);
          r.forward = f.logLike();
          r.backward = b.logLike();
          threadCounts[tid].loglike += r.forward;
        } else if (doBackward) { const BackwardMatrix b (eval, sp); r.backward = b.logLike(); }
        if (doPostTrans) {
          // BackwardMatrix::postTransQueue (backward.cpp:52-56): every (cell, transition) posterior, largest first; then
          // BackwardMatrix::traceFrom with a TraceTerminator (backward.cpp:98-108) from the first entry's SOURCE cell
          const ForwardMatrix f (eval, sp);
          const BackwardMatrix b (eval, sp);
          r.forward = f.logLike();
          ostringstream m;
          if (r.forward > -numeric_limits<double>::infinity()) {
            BackwardMatrix::PostTransQueue q = b.postTransQueue (f);
            m << ",\"postTransCount\":" << q.size() << ",\"postTrans\":[";
            BackwardMatrix::PostTrans first = q.top();
            for (int n = 0; n < 12 && !q.empty(); ++n) {
              const BackwardMatrix::PostTrans pt = q.top();
              q.pop();
              m << (n ? "," : "") << "[" << pt.inPos << "," << pt.outPos << "," << (eval.state[pt.src].transOffset + pt.transIndex) << "," << dstr (pt.weight) << "]";
            }
            m << "]";
            const MachineTransition& mt = machine.state[first.src].getTransition (first.transIndex);
            const long srcIn = first.inPos - (mt.inputEmpty() ? 0 : 1), srcOut = first.outPos - (mt.outputEmpty() ? 0 : 1);
            vector<long long> visited;
            BackwardMatrix::TraceTerminator collect = [&] (Envelope::InputIndex, Envelope::OutputIndex, StateIndex src, EvaluatedMachineState::TransIndex ti) {
              visited.push_back ((long long) (eval.state[src].transOffset + ti));
              return false;
            };
            b.traceFrom (machine, f, srcIn, srcOut, first.src, first.transIndex, collect);
            m << ",\"traceFrom\":[";
            for (size_t n = 0; n < visited.size(); ++n) m << (n ? "," : "") << visited[n];
            m << "]";
          }
          r.postTrans = m.str();
        }
        if (doMatrices) {
          const ForwardMatrix f (eval, sp);
          const BackwardMatrix b (eval, sp);
          const ViterbiMatrix v (eval, sp);
          r.forward = f.logLike(); r.backward = b.logLike(); r.viterbi = v.logLike();
          // layout [o][i][s], the reference's storage order (dpmatrix.h:86-95)
          ostringstream m;
          const char* nm[3] = { "F", "B", "V" };
          for (int which = 0; which < 3; ++which) {
            m << ",\"" << nm[which] << "\":[";
            size_t n = 0;
            for (long o = 0; o <= f.outLen; ++o)
              for (long i = 0; i <= f.inLen; ++i)
                for (StateIndex s = 0; s < f.nStates; ++s)
                  m << (n++ ? "," : "") << dstr (which == 0 ? f.cell (i, o, s) : which == 1 ? b.cell (i, o, s) : v.cell (i, o, s));
            m << "]";
          }
          r.matrices = m.str();
        }
      }
    };

    const auto t0 = chrono::steady_clock::now();
    if (nThreads <= 1) work (0);
    else {
      vector<thread> th;
      for (int t = 0; t < nThreads; ++t) th.emplace_back (work, t);
      for (auto& t: th) t.join();
    }
    const double secs = chrono::duration<double> (chrono::steady_clock::now() - t0).count();

    MachineCounts total (eval);
    double cs = 0;
    for (int t = 0; t < nThreads; ++t) { total += threadCounts[t]; total.loglike += threadCounts[t].loglike; cs += cellStates[t]; }

    cout << "{\"seconds\":" << dstr (secs) << ",\"threads\":" << nThreads << ",\"nPairs\":" << pairs.size()
         << ",\"cellStatesPerPass\":" << dstr (cs) << ",\"nStates\":" << eval.nStates() << ",\"nTrans\":" << eval.nTransitions;
    if (!quiet) {
      cout << ",\n \"pairs\":[";
      for (size_t k = 0; k < pairs.size(); ++k) {
        const PairResult& r = res[k];
        cout << (k ? ",\n  " : "\n  ") << "{\"tokenizable\":" << (r.tokenizable ? "true" : "false");
        const double ninf = -numeric_limits<double>::infinity();
        if (doRolling) cout << ",\"rolling\":" << dstr (r.tokenizable ? r.rolling : ninf);
        if (doForward || doCounts || doMatrices || doSample || doPostTrans) cout << ",\"forward\":" << dstr (r.tokenizable ? r.forward : ninf);
        if (doBackward || doCounts || doMatrices) cout << ",\"backward\":" << dstr (r.tokenizable ? r.backward : ninf);
        if (doViterbi || doPath || doMatrices) cout << ",\"viterbi\":" << dstr (r.tokenizable ? r.viterbi : ninf);
        if (doPath) {
          cout << ",\"path\":[";
          for (size_t n = 0; n < r.path.size(); ++n) cout << (n ? "," : "") << r.path[n];
          cout << "]";
        }
        if (doSample) {
          cout << ",\"sample\":[";
          for (size_t n = 0; n < r.sample.size(); ++n) cout << (n ? "," : "") << r.sample[n];
          cout << "]";
        }
        if (pairs[k].alignment.size()) {   // the path envelope every matrix of this pair was given (seqpair.cpp:104-110)
          const Envelope env (pairs[k]);
          cout << ",\"env\":";
          env.writeJson (cout);
        }
        cout << r.matrices << r.postTrans << "}";
      }
      cout << "\n ]";
    }
    if (doCounts) {
      cout << ",\n \"loglike\":" << dstr (total.loglike) << ",\n \"counts\":[";
      size_t n = 0;
      for (const auto& cv: total.count) for (double c: cv) cout << (n++ ? "," : "") << dstr (c);
      cout << "]";
    }
    cout << "}" << endl;
  } catch (const std::exception& e) {
    cerr << "refdrv: " << e.what() << endl;
    return 1;
  }
  return 0;
}
