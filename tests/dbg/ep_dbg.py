import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/dsp-speech-recognition_b200')
import numpy as np, dspfe
from dspfe import synth
pcm, off = synth.synth_batch([8000, 12345, 16000], seed0=1)
plan = dspfe.EndpointPlan()
print(plan.detect_host(pcm, off))
