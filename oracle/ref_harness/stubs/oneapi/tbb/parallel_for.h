/* empty stub: the reference includes this header but never uses it (PF/apps/rayleighTaylor2D.h:19-20) */
