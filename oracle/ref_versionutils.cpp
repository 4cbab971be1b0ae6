// oracle/ref_versionutils.cpp -- TEST INFRASTRUCTURE (part of the oracle/_ref build recipe).
//
// The reference generates versionutils.cpp from versionutils.cpp.in with CMake's
// configure_file (CMakeLists.txt:155-158); that step writes into the source tree, which is
// read-only here. This file is our own definition of the symbols declared in the
// reference's versionutils.h so the unmodified engine links. Version 1.8.5 is what
// CMakeLists.txt:53-55 sets (the addon refuses a mismatching engine, bake.py:3247-3251).
#include "versionutils.h"

int VersionUtils::_major = 1;
int VersionUtils::_minor = 8;
int VersionUtils::_revision = 5;
std::string VersionUtils::_label = "1.8.5 ffb200 reference build";
std::string VersionUtils::_support_license_type = "GitHub";
std::string VersionUtils::_support_license_id = "000AA";

void VersionUtils::getVersion(int *major, int *minor, int *revision) {
    *major = _major; *minor = _minor; *revision = _revision;
}
int VersionUtils::getMajor() { return _major; }
int VersionUtils::getMinor() { return _minor; }
int VersionUtils::getRevision() { return _revision; }
std::string VersionUtils::getLabel() { return _label; }
std::string VersionUtils::getSupportLabel() { return _support_license_type + " " + _support_license_id; }
std::string VersionUtils::getSupportLicenseID() { return _support_license_id; }
