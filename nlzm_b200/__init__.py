"""B200-native match finding for NLZM, and the host pipeline around it.

  matchfinder.MatchFinders      the reference's finder verbs (Init / FindAndUpdate / Shift / Release)
                                over libnlzm_mf (hand-written sm_100a CUDA behind include/nlzm_mf.h)
  codec.compress / decompress   libnlzm_codec: own parser, nibble model and rANS frame coder over the
                                engine; the reference's stream format, byte for byte
  sharding                      position ranges per GPU (no collective on the data path)
  synth                         the seeded synthetic workloads of BASELINE.json
  build                         in-tree builds (nvcc for the engine, g++ for the host pipeline)

Nothing here falls back to a CPU matcher: without the CUDA library or a device, calls raise.
"""
