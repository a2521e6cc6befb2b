# the whole GPU suite, smoke, then the short bench
( time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
NOTEST=1 bash profiles/tools/r4_check.sh
