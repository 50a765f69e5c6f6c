// Legacy-ASCII VTK output and input, and the progress line (reference:
// include/vtk.cuh; format: http://www.vtk.org/wp-content/uploads/2015/04/
// file-formats.pdf).
//
// Vtk_output writes one file per call of write_positions,
// <output_path><base_name>_<frame>.vtk, and appends the further sections
// (links, fields, polarities, properties) to that file. Vtk_input reads the
// sections back by keyword. Host-only; not part of the timed step.
#pragma once

#include <assert.h>
#include <sys/stat.h>
#include <time.h>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <typeinfo>
#include <vector>

#include "links.cuh"
#include "polarity.cuh"
#include "utils.cuh"


template<typename Pt, template<typename> class Solver>
class Solution;

template<typename Prop>
struct Property;


class Vtk_output {
public:
    // Files are stored as output_path/base_name_#.vtk
    Vtk_output(std::string base_name, std::string output_path = "output/",
        bool verbose = true);
    ~Vtk_output(void);
    // Write x, y, and z component of Pt; has to be written first. Points with
    // input_mask[i] == false are left out of this and all following sections.
    template<typename Pt, template<typename> class Solver>
    void write_positions(Solution<Pt, Solver>& points, bool* input_mask = NULL);
    // Write links, see links.cuh; if written has to be second
    void write_links(Links& links);
    // Write further components of Pt
    template<typename Pt, template<typename> class Solver>
    void write_field(Solution<Pt, Solver>& points, const char* data_name = "w",
        float Pt::*field = &Pt::w);
    // Write a polarity of Pt as unit normals, see polarity.cuh; the default
    // theta = phi = 0 is written as {0, 0, 0}.
    template<typename Pt, float Pt::*theta = &Pt::theta,
        float Pt::*phi = &Pt::phi, template<typename> class Solver>
    void write_polarity(
        Solution<Pt, Solver>& points, const char* data_name = "polarity");
    // Write not integrated property, see property.cuh
    template<typename Prop>
    void write_property(Property<Prop>& property);

private:
    int n_points = 0;
    int n_to_write = 0;
    bool* mask = NULL;
    int time_step = 0;
    std::string base_name;
    std::string output_dir;
    std::string current_path;
    bool verbose;
    bool point_data_started = false;
    time_t t_0;

    bool skipped(int i) const { return mask != NULL && !mask[i]; }

    // Re-open the current frame for appending; the first POINT_DATA section
    // writer also emits the section header.
    std::ofstream append_point_data()
    {
        std::ofstream file(current_path, std::ios_base::app);
        assert(file.is_open());
        if (!point_data_started) {
            file << "\nPOINT_DATA " << n_to_write << "\n";
            point_data_started = true;
        }
        return file;
    }
};

inline Vtk_output::Vtk_output(
    std::string base_name, std::string output_path, bool verbose)
    : base_name{base_name}, output_dir{output_path}, verbose{verbose}
{
    if (output_dir.empty() || output_dir.back() != '/') {
        output_dir.append("/");
        std::cout << output_dir << std::endl;
    }
    mkdir(output_dir.c_str(), 0755);
    time(&t_0);
}

inline Vtk_output::~Vtk_output()
{
    if (!verbose) return;

    const auto duration = time(NULL) - t_0;
    std::cout << "Integrating " << base_name << ", ";
    if (duration < 60)
        std::cout << duration << " seconds";
    else if (duration < 60 * 60)
        std::cout << duration / 60 << "m " << duration % 60 << "s";
    else
        std::cout << duration / (60 * 60) << "h " << duration % (60 * 60)
                  << "m";
    std::cout << " taken (" << n_points << " points).        \n";
}

template<typename Pt, template<typename> class Solver>
void Vtk_output::write_positions(Solution<Pt, Solver>& points, bool* input_mask)
{
    n_points = *points.h_n;
    mask = input_mask;
    n_to_write = 0;
    for (int i = 0; i < n_points; i++) n_to_write += skipped(i) ? 0 : 1;

    current_path =
        output_dir + base_name + "_" + std::to_string(time_step) + ".vtk";
    std::ofstream file(current_path);
    assert(file.is_open());

    file << "# vtk DataFile Version 3.0\n"
         << base_name << "\n"
         << "ASCII\n"
         << "DATASET POLYDATA\n"
         << "\nPOINTS " << n_to_write << " float\n";
    for (int i = 0; i < n_points; i++) {
        if (skipped(i)) continue;
        const Pt& X = points.h_X[i];
        file << X.x << " " << X.y << " " << X.z << "\n";
    }

    file << "\nVERTICES " << n_to_write << " " << 2 * n_to_write << "\n";
    for (int i = 0; i < n_to_write; i++) file << "1 " << i << "\n";

    point_data_started = false;
    time_step += 1;
    if (!verbose) return;

    std::cout << "Integrating " << base_name << ", " << time_step
              << " steps done (" << n_points << " points)        \r";
    std::cout.flush();
}

inline void Vtk_output::write_links(Links& links)
{
    std::ofstream file(current_path, std::ios_base::app);
    assert(file.is_open());

    const int n_links = *links.h_n;
    file << "\nLINES " << n_links << " " << 3 * n_links << "\n";
    for (int i = 0; i < n_links; i++)
        file << "2 " << links.h_link[i].a << " " << links.h_link[i].b << "\n";
}

template<typename Pt, template<typename> class Solver>
void Vtk_output::write_field(
    Solution<Pt, Solver>& points, const char* data_name, float Pt::*field)
{
    std::ofstream file = append_point_data();
    file << "SCALARS " << data_name << " float\n"
         << "LOOKUP_TABLE default\n";
    for (int i = 0; i < n_points; i++) {
        if (skipped(i)) continue;
        file << points.h_X[i].*field << "\n";
    }
}

template<typename Pt, float Pt::*theta, float Pt::*phi,
    template<typename> class Solver>
void Vtk_output::write_polarity(
    Solution<Pt, Solver>& points, const char* data_name)
{
    std::ofstream file = append_point_data();
    file << "NORMALS " << data_name << " float\n";
    for (int i = 0; i < n_points; i++) {
        if (skipped(i)) continue;
        const Pt& X = points.h_X[i];
        float3 n = pol_to_float3<Pt, theta, phi>(X);
        if (X.*theta == 0 && X.*phi == 0) n.z = 0;  // "no polarity"
        file << n.x << " " << n.y << " " << n.z << "\n";
    }
}

template<typename Prop>
void Vtk_output::write_property(Property<Prop>& property)
{
    std::ofstream file = append_point_data();
    // float properties are written as floats, everything else as int
    const std::string type_name =
        std::string(typeid(Prop).name()) == "f" ? "float" : "int";

    assert(n_points <= property.n_max);
    file << "SCALARS " << property.name << " " << type_name << "\n"
         << "LOOKUP_TABLE default\n";
    for (int i = 0; i < n_points; i++) {
        if (skipped(i)) continue;
        file << property.h_prop[i] << "\n";
    }
}


class Vtk_input {
public:
    Vtk_input(std::string file_name);
    // Stream position just behind the line that starts with the two keywords
    std::streampos find_entry(std::string, std::string);
    template<typename Pt, template<typename> class Solver>
    void read_positions(Solution<Pt, Solver>& points);
    // Read polarity of Pt, see polarity.cuh
    template<typename Pt, template<typename> class Solver>
    void read_polarity(Solution<Pt, Solver>& points);
    // Read further field of Pt
    template<typename Pt, template<typename> class Solver>
    void read_field(Solution<Pt, Solver>& points, const char* data_name = "w",
        float Pt::*field = &Pt::w);
    // Read property, see property.cuh
    template<typename Prop>
    void read_property(Property<Prop>& property, std::string prop_name);
    int n_points;

private:
    std::string file_name;

    // Stream positioned on the first data line of the section.
    std::ifstream open_at(std::string keyword1, std::string keyword2,
        int lines_to_skip = 0)
    {
        const std::streampos where = find_entry(keyword1, keyword2);
        std::ifstream input_file(file_name);
        assert(input_file.is_open());
        input_file.seekg(where);
        std::string line;
        for (int i = 0; i < lines_to_skip; i++) getline(input_file, line);
        return input_file;
    }
};

inline Vtk_input::Vtk_input(std::string file_name) : file_name{file_name}
{
    std::ifstream input_file(file_name);
    assert(input_file.is_open());

    // "POINTS <n> float" is on one of the first six lines
    n_points = 0;
    std::string line;
    for (int i = 0; i < 6; i++) {
        getline(input_file, line);
        const auto items = split(line);
        if (items.size() > 1 && items[0] == "POINTS") {
            n_points = stoi(items[1]);
            break;
        }
    }
}

inline std::streampos Vtk_input::find_entry(
    std::string keyword1, std::string keyword2)
{
    std::ifstream input_file(file_name);
    assert(input_file.is_open());

    std::string line;
    for (int i = 0; i < 4; i++) getline(input_file, line);  // header

    while (getline(input_file, line)) {
        const auto items = split(line);
        if (items.size() > 1 && items[0] == keyword1 && items[1] == keyword2)
            return input_file.tellg();
    }
    std::cout << "Vtk_input: no entry \"" << keyword1 << " " << keyword2
              << "\" in " << file_name << std::endl;
    assert(false);
    return input_file.tellg();
}

template<typename Pt, template<typename> class Solver>
void Vtk_input::read_positions(Solution<Pt, Solver>& points)
{
    std::ifstream input_file = open_at("POINTS", std::to_string(n_points));
    std::string line;
    for (int i = 0; i < n_points; i++) {
        getline(input_file, line);
        const auto items = split(line);
        points.h_X[i].x = stof(items[0]);
        points.h_X[i].y = stof(items[1]);
        points.h_X[i].z = stof(items[2]);
    }
}

template<typename Pt, template<typename> class Solver>
void Vtk_input::read_polarity(Solution<Pt, Solver>& points)
{
    std::ifstream input_file = open_at("NORMALS", "polarity");
    std::string line;
    for (int i = 0; i < n_points; i++) {
        getline(input_file, line);
        const auto items = split(line);
        const auto x = stof(items[0]);
        const auto y = stof(items[1]);
        const auto z = stof(items[2]);
        const auto dist = sqrt(pow(x, 2) + pow(y, 2) + pow(z, 2));
        if (dist == 0) {  // written for theta = phi = 0
            points.h_X[i].phi = 0.0f;
            points.h_X[i].theta = 0.0f;
        } else {
            points.h_X[i].phi = atan2(y, x);
            points.h_X[i].theta = acos(z);  // the normals are unit vectors
        }
    }
}

template<typename Pt, template<typename> class Solver>
void Vtk_input::read_field(
    Solution<Pt, Solver>& points, const char* data_name, float Pt::*field)
{
    std::ifstream input_file = open_at("SCALARS", data_name, 1);  // LOOKUP_TABLE
    std::string line;
    for (int i = 0; i < n_points; i++) {
        getline(input_file, line);
        std::istringstream(line) >> points.h_X[i].*field;
    }
}

template<typename Prop>
void Vtk_input::read_property(Property<Prop>& property, std::string prop_name)
{
    std::ifstream input_file = open_at("SCALARS", prop_name, 1);  // LOOKUP_TABLE
    assert(n_points <= property.n_max);
    std::string line;
    for (int i = 0; i < n_points; i++) {
        getline(input_file, line);
        std::istringstream(line) >> property.h_prop[i];
    }
}


// Extension: Vtk_async_output<Pt>, snapshots on the solver's stream and writes
// frames (ASCII or binary) from a background thread.
#include "b200/vtk_async.cuh"
