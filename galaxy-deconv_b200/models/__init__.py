"""Drop-in replacements for the reference's ``models`` package on the inference hot path (same module paths, class
names, constructor arguments, ``forward`` signatures and ``state_dict`` layouts), executing on libgdeconv's
sm_100a kernels.  Put ``galaxy-deconv_b200/`` on ``sys.path`` (ahead of the reference checkout) and ``test.py`` /
``test_psf.py`` / the tutorials import these instead."""
