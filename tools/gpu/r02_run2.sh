# round 2, run 2: the whole GPU suite (new: cs32 demodulator, audio compat library, extensions, comm gather, b200_multi)
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest_run2.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02/pytest_run2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
