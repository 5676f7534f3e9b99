// COMPILE-CHECK STAND-IN for the reference's radio.h (the Qt main window): the adapter only connects its
// signals to RadioInterface's slots by name
#pragma once
#include "QObject"
class RadioInterface : public QObject {};
