"""Host emulation of the device logic -- TEST INFRASTRUCTURE ONLY (see emu_device.h)."""
