// FasTC::Pixel, 8 bits per channel.  The reference's Pixel carries a per-channel bit depth
// (Base/include/FasTC/Pixel.h); the compression path only ever sees RGBA8 through
// Pack()/Unpack() (Base/src/Pixel.cpp:165-179: R in the low byte), which is what this keeps.
#ifndef FASTC_B200_PIXEL_H_
#define FASTC_B200_PIXEL_H_
#include "FasTC/TexCompTypes.h"

namespace FasTC {

class Pixel {
 public:
  Pixel() : m_R(0), m_G(0), m_B(0), m_A(0) {}
  Pixel(uint8 a, uint8 r, uint8 g, uint8 b) : m_R(r), m_G(g), m_B(b), m_A(a) {}
  explicit Pixel(uint32 rgba) { Unpack(rgba); }

  uint8 &R() { return m_R; }
  uint8 &G() { return m_G; }
  uint8 &B() { return m_B; }
  uint8 &A() { return m_A; }
  const uint8 &R() const { return m_R; }
  const uint8 &G() const { return m_G; }
  const uint8 &B() const { return m_B; }
  const uint8 &A() const { return m_A; }

  // R | G << 8 | B << 16 | A << 24
  uint32 Pack() const { return (uint32)m_R | ((uint32)m_G << 8) | ((uint32)m_B << 16) | ((uint32)m_A << 24); }
  void Unpack(uint32 rgba) {
    m_R = rgba & 0xFF; m_G = (rgba >> 8) & 0xFF; m_B = (rgba >> 16) & 0xFF; m_A = rgba >> 24;
  }
  void MakeOpaque() { m_A = 255; }
  bool operator==(const Pixel &o) const { return Pack() == o.Pack(); }
  bool operator!=(const Pixel &o) const { return Pack() != o.Pack(); }

 private:
  uint8 m_R, m_G, m_B, m_A;
};

}  // namespace FasTC
#endif
