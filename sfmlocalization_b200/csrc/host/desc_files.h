// desc_files.h -- on-disk formats either side of the matching stage:
// .desc descriptor files (FileUtils.cpp:77-103), OpenMVG match text files (FileUtils.cpp:106-120),
// pair list files (FileUtils.cpp:180-...), and the view list of sfm_data.json.
#pragma once
#include <string>
#include <vector>

#include "hulo_types.h"

namespace hulo {

// hulo::readAKAZEBin, FileUtils.cpp:94-103: uint64 LE count, then count x 64 raw bytes.
// rows receives count x 64 bytes.  Returns false when the file cannot be read (the reference
// then matches against an empty matrix).
bool readAKAZEBin(const std::string &filename, std::vector<uint8_t> &rows, std::size_t &count);
// hulo::saveAKAZEBin, FileUtils.cpp:77-92: rows of `width` bytes (61 for AKAZE MLDB) are zero
// padded to 64.
bool saveAKAZEBin(const std::string &filename, const uint8_t *rows, std::size_t count, std::size_t width);

// hulo::exportPairWiseMatches, FileUtils.cpp:106-120 (openMVG::matching::Save, text flavour):
// per pair "I J\nN\n" followed by N lines "i j".
bool exportPairWiseMatches(const PairWiseMatches &matches, const std::string &filename);
bool importPairWiseMatches(const std::string &filename, PairWiseMatches &matches);

// hulo::readPairFile: whitespace separated "I J" per line.
bool readPairFile(const std::string &filename, std::vector<Pair> &pairs);

// The view list (id_view, filename) of an OpenMVG cereal sfm_data.json, which is all the
// matchers use of SfM_Data (MatchUtils.cpp:86-91, 328-330).
bool readViewsFromSfmData(const std::string &sfm_data_json, Views &views);
// Plain alternative: one "id path" per line.
bool readViewsFromList(const std::string &filename, Views &views);

// OpenMVG .feat file of a view (Regions::Load, text flavour): one feature per line,
// "x y scale orientation" for the SIOPointFeature of AKAZE regions ("x y" alone is accepted).
bool readFeatFile(const std::string &filename, FeatureLocations &feats);
// Regions_Provider::load for the views given: <dir>/<basename>.feat for each.
bool loadRegions(const Views &views, const std::string &dir, RegionsProvider &regions);
std::string featPath(const std::string &dir, const std::string &img_path);

// hulo::readMatBin / saveMatBin, FileUtils.cpp:44-75: int32 rows, cols, OpenCV type code, then the
// raw elements (depths CV_8U 0, CV_32S 4, CV_32F 5, CV_64F 6; one channel).  Values come back as
// doubles, row-major.  Used for the <view>.bow bag-of-features vectors (BoFUtils.cpp:38-42).
bool readMatBin(const std::string &filename, int &rows, int &cols, std::vector<double> &values);
bool saveMatBin(const std::string &filename, int rows, int cols, int cv_type, const double *values);
std::string bowPath(const std::string &dir, const std::string &img_path);

// stlplus::create_filespec(dir, basename_part(path), "desc")
std::string descPath(const std::string &dir, const std::string &img_path, bool strip_extension);

}  // namespace hulo
