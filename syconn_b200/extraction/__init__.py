"""Drop-in mirrors of ``syconn.extraction``'s hot-path modules (same names, signatures and outputs)."""
