"""CPU oracle for the xrdslam render-and-optimise hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``xrdslam_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
as the timed CPU baseline -- never as the product path.

Contents
  tcnn_restated.py   torch-CPU restatement of tinycudann HashGrid / OneBlob
                     (third-party, NOT vendored in the reference: parity UNPINNED
                     against real tcnn -- see DESIGN.md section "Oracle").
  coslam.py          restatement of slam/models/joint_encoding.py (+ utils.py losses)
  nice.py            restatement of slam/models/conv_onet.py + decoder_nice.py
  voxfusion.py       restatement of sparse_voxel.py + the svo/grid native code
  pointslam.py       restatement of conv_onet_pointslam.py + exact kNN
  ref_harness.py     imports the REAL reference classes from /root/reference with
                     sys.modules stubs (only available in the build container)
  synthetic.py       the synthetic 640x480 RGB-D room used by tests and bench
"""
