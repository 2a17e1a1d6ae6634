# scratch: the command file handed to gpurun by profiles/gpurun_retry.sh (last used for the final validation of round 2)
python -m pytest tests -q -m gpu
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py
