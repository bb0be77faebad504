from .window import assemble_windows, frame_window, get_affine_transform  # noqa: F401
