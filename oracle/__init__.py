"""CPU oracle for the predict + Grad-CAM hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker / CPU baseline.

Pinning status
--------------
* ``oracle.cnn``  (both CNN flavours, forward + explain-backward): PINNED to
  outputs of the reference's own code run in the authoring container
  (``tests/golden/make_golden.py`` imports ``Classes/CNNModel.py``,
  ``WebApplicationPrototype/explainability.py`` and
  ``WebApplicationPrototype/ADCNNM.py`` from ``/root/reference`` and commits
  their outputs as ``tests/golden/*.npz``).
* ``oracle.gradcam`` (Grad-CAM tail / overlay): the arithmetic lives in the
  third-party package ``pytorch-grad-cam`` (PyPI ``grad-cam``, un-pinned by the
  reference, absent from ``/root/reference`` and from this image).  The tail is
  a restatement of its published algorithm, anchored on the reference's call
  sites ``WebApplicationPrototype/GRADCAM.py:53,64,67,70``.  The reference
  holds no golden vectors for it => **parity unpinned** for the composition;
  the sub-steps that call OpenCV (``cv2.resize`` bilinear, ``COLORMAP_JET``)
  ARE pinned against the OpenCV build the reference calls
  (``tests/golden/cv2_resize.npz``, ``tests/golden/jet_lut.npy``).
"""
