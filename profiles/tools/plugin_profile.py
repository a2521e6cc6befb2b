import cProfile, pstats, sys, os
sys.argv = ['x', '300000']
sys.path.insert(0, '/root/repo')
cProfile.run(open('/root/repo/profiles/tools/plugin_throughput.py').read(), 'gpurun_out/pp.out')
pstats.Stats('gpurun_out/pp.out').sort_stats('cumtime').print_stats(18)
