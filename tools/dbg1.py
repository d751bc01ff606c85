import importlib, sys, numpy as np
sys.path.insert(0, '.')
p = importlib.import_module('jsvm-code-notes_b200'); s = importlib.import_module('jsvm-code-notes_b200.synth')
me = p.MotionEstimator(0)
W, H = 176, 144
me.configure(W, H, 2); me.sync(); print('configured')
me.set_params(16, 2)
seq = s.Sequence(W, H, 5)
me.upload_plane(0, seq.frame(0)); me.sync(); print('uploaded 0')
me.upload_plane(1, seq.frame(1)); me.sync(); print('uploaded 1')
me.interp_halfpel([0]); me.sync(); print('interp ok')
hp = me.halfpel_plane(0); print(hp.shape, hp[112:116, 112:120])
