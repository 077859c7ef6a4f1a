// fc_tune.h — timing-experiment knobs. The shipped library is built WITHOUT FC_TUNING: every knob is then its
// compile-time default and no environment variable is read anywhere in the product path. A development build
// (FFTCONV_B200_TUNING=1 python -m fft_conv_pytorch_b200.build) reads FFTCONV_B200_<NAME> once per process for A/B runs.
#pragma once
#include <cstdlib>
#include <cstring>

#ifdef FC_TUNING
inline const char* fc_tune_str(const char* name) {
  char buf[96] = "FFTCONV_B200_";
  std::strncat(buf, name, sizeof(buf) - std::strlen(buf) - 1);
  return std::getenv(buf);
}
inline int fc_tune_int(const char* name, int dflt) {
  const char* e = fc_tune_str(name);
  return e ? std::atoi(e) : dflt;
}
#else
inline const char* fc_tune_str(const char*) { return nullptr; }
inline int fc_tune_int(const char*, int dflt) { return dflt; }
#endif
