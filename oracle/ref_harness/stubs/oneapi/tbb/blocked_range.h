/* empty stub: the reference includes this header but never uses it (PF/apps/laplace3D.h:19-20) */
