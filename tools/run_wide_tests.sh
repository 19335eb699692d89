# run under gpurun: the wide engine's parity tests one by one (each under its own timeout), then its throughput
cd $GRAFT_REPO_ROOT
for k in "prot2dna_dnapsw-2" "wide" "2] or -2"; do
  echo "== $k"; timeout 200 python -m pytest tests -m gpu -x -q -k "$k" 2>&1 | tail -3
done
for g in 32 8; do echo "== G=$g"; MB_WIDE_G=$g timeout 200 python -m pytest tests -m gpu -x -q -k "wide or prot2dna_dnapsw-2 or ragged_batch or hmmer_pf00516-2" 2>&1 | tail -2; done
MB_WIDE_VERBOSE=1 timeout 120 python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 148 --li 300 --lo 10000 --engines 2 2>&1 | tail -2
for g in 32 8; do MB_WIDE_G=$g MB_WIDE_VERBOSE=1 timeout 60 python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 148 --li 300 --lo 2000 --engines 2 --no-trace --reps 1 2>&1 | tail -2; done
MB_WIDE_VERBOSE=1 timeout 60 python tools/bench_wide.py --machine hmmer_pf00516 --pairs 2048 --li 0 --lo 275 --engines 2 2>&1 | tail -2
