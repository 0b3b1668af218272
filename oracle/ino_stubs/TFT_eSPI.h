// TEST INFRASTRUCTURE stand-in: see arduino_stub.h
#include "arduino_stub.h"
