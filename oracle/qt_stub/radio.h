// TEST INFRASTRUCTURE: the reference's rds-decoder-1.cpp includes "radio.h" only to know the
// RadioInterface type it stores a pointer to (never dereferenced).  This stand-in is found first
// on the include path of oracle/Makefile; the real radio.h (Qt GUI) is not used.
#pragma once
class RadioInterface {};
