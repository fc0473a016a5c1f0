import sys, os, cProfile, pstats
sys.argv = ["sweep_check.py", "256", "64"]
pr = cProfile.Profile()
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sweep_check.py")).read()
src = src.replace("sb.spectrum_matter_sweep(prob, names, th[:8], ks)  # warm-up", "sb.spectrum_matter_sweep(prob, names, th[:8], ks); pr.enable()  #")
src = src.replace('upd = sb.parameter_updater(prob, names)\nfor i in', 'pr.disable(); upd = sb.parameter_updater(prob, names)\nfor i in')
exec(compile(src, "sweep_check.py", "exec"), dict(__name__="__main__", pr=pr, __file__=os.path.join(os.path.dirname(os.path.abspath(__file__)), "sweep_check.py")))
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
