"""CPU oracle for the InsMOS sparse-voxel forward path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``insmos_b200/`` imports this package.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline legs may import, link or execute anything under ``oracle/`` -- and
only as the checker / the timed CPU baseline, never as the thing shipped.

What it restates (paths relative to the reference repository, nubot-nudt/InsMOS):

* ``oracle/me.py``      MinkowskiEngine CPU semantics used by models/backbones_3d/motionnet.py:21-50
                        and models/MinkowskiEngine/minkunet.py:52-181 (sparse_collate, TensorField.sparse,
                        coordinate stride, kernel maps, gather -> sgemm -> scatter-add convolution,
                        transposed convolution, slice).
* ``oracle/sp.py``      spconv 2.3.6 semantics used by models/backbones_3d/voxel_generate.py:17-31 and
                        models/backbones_3d/spconv_unet.py:120-410 (PointToVoxel.generate_voxel_with_id,
                        SubM / strided / inverse convolution with indice_key reuse, dense(), gather by id).
* ``oracle/native/oracle_native.c``  kernel-map neighbour table (hash map of the input coordinates, kernel offsets probed in
                        parallel with OpenMP -- how the two libraries' CPU backends build their maps; checked against the
                        numpy statement in me.py / sp.py), rotated-BEV IoU + NMS (models/bbox_post_process/src/iou3d_nms_kernel.cu:15-311,
                        host twin iou3d_cpu.cpp:38-228, sweep iou3d_nms.cpp:90-136) and
                        Array_Index.find_features_by_bbox_with_yaw (models/utils/src/Array_Index.cpp:14-79).
* ``oracle/graph.py``   the model graph models/models.py:297-376 + the modules it calls, as one
                        functional forward over a state_dict.
* ``oracle/shims/``     oracle-backed stand-ins for the external packages (MinkowskiEngine, spconv,
                        pytorch_lightning) so that the reference's OWN python model code can be
                        imported in the development container to generate tests/golden fixtures.

PARITY STATUS:
* MinkowskiEngine and spconv are external, un-vendored dependencies of the reference (ME: un-pinned
  NVIDIA/MinkowskiEngine master ~v0.5.4; spconv: spconv_cu113==2.3.6).  Neither is installed nor
  present in source form, and the reference has no tests or golden vectors.  The restatement of those
  two libraries' operators follows their published algorithms and is anchored
  on the reference's call sites; it is self-validated against dense torch convolutions and algebraic
  properties (tests/test_oracle_*.py) -- **parity unpinned** at the ME/spconv operator boundary.
* The first-party native code IS pinned: ``oracle/build_ref.py`` compiles the reference's own
  Array_Index.cpp and iou3d_cpu.cpp / iou3d_nms(.cpp/.cu) from where they lie into ``oracle/_ref/``
  and tests/test_oracle_native.py checks the C restatement against them bit for bit.
* The model graph is pinned by running the reference's own models/*.py over the oracle shims
  (tests/golden/make_golden.py) and comparing ``oracle/graph.py`` and the CUDA path with the result.
"""
