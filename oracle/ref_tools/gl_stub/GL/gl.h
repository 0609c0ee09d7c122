/* empty stand-in for <GL/gl.h>: the reference's headers include <cuda_gl_interop.h>; glad.h already defines the GL types */
