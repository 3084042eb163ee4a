"""CPU oracle for the WhisperLive per-chunk hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``whisperlive_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it, and then
only as the checker / CPU baseline, never as the thing measured or shipped.

Pinning status (see DESIGN.md "Oracle"):
  * mel            -- pinned against the reference's in-repo mel formula
                      (whisper_live/transcriber/tensorrt_utils.py:130-194, executed
                      here with stubbed I/O imports) and HF transformers'
                      WhisperFeatureExtractor; fixtures in tests/golden/.
  * network math   -- pinned against HF transformers 5.5.0 modeling_whisper.py
                      (random weights); fixtures in tests/golden/.
  * host logic     -- pinned against the reference's own
                      transcriber_faster_whisper.py executed with stubbed
                      ctranslate2/faster_whisper imports; fixtures in tests/golden/.
  * CT2 search / timestamp rules / align -- **parity unpinned** at the CTranslate2
                      boundary: CTranslate2 and faster-whisper are not installed and
                      not vendored in /root/reference; restated from their published
                      algorithm.  Closest executable anchors (tests/test_oracle_golden.py):
                      the logits rules (suppress list, suppress-blank, timestamp rules,
                      log-softmax), the median filter, DTW and the alignment pipeline
                      order agree exactly with transformers' ports of OpenAI whisper's
                      ApplyTimestampRules / timing.py -- the code CTranslate2 itself ports.
                      The beam search proper (CT2 decoding.cc) has no such anchor.
"""
