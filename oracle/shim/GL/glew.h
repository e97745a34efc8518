// oracle shim (test infrastructure): GL typedefs/enums and no-op entry points referenced by
// src/gl/Texture.{hpp,cpp} and src/Util.cpp:81-97.  Nothing on the oracle path ever calls them.
#pragma once
#include <cstdint>
typedef unsigned int GLuint;
typedef int GLint;
typedef int GLsizei;
typedef unsigned int GLenum;
typedef std::uint64_t GLuint64;
enum : GLenum {
  GL_LINEAR = 0x2601, GL_NEAREST = 0x2600, GL_REPEAT = 0x2901, GL_RGBA8 = 0x8058, GL_RGB = 0x1907, GL_RGBA = 0x1908,
  GL_UNSIGNED_BYTE = 0x1401, GL_TEXTURE_2D = 0x0DE1, GL_TEXTURE_WRAP_S = 0x2802, GL_TEXTURE_WRAP_T = 0x2803,
  GL_TEXTURE_MIN_FILTER = 0x2801, GL_TEXTURE_MAG_FILTER = 0x2800, GL_TEXTURE_WIDTH = 0x1000, GL_TEXTURE_HEIGHT = 0x1001
};
inline void glDeleteTextures(GLsizei, const GLuint*) {}
inline void glCreateTextures(GLenum, GLsizei, GLuint* ids) { if (ids) *ids = 0; }
inline void glTextureStorage2D(GLuint, GLsizei, GLenum, GLsizei, GLsizei) {}
inline void glTextureParameteri(GLuint, GLenum, GLint) {}
inline void glBindTextureUnit(GLuint, GLuint) {}
inline void glMakeTextureHandleNonResidentARB(GLuint64) {}
inline void glMakeTextureHandleResidentARB(GLuint64) {}
inline void glGetTextureLevelParameteriv(GLuint, GLint, GLenum, GLint* p) { if (p) *p = 0; }
inline void glGetTextureImage(GLuint, GLint, GLenum, GLenum, GLsizei, void*) {}
