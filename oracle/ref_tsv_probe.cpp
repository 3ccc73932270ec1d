// TEST INFRASTRUCTURE (built by oracle/Makefile.ref into oracle/_ref/tsv_probe): reads an assignment TSV the way the
// reference's downstream programs do (hmmufotu-sum, src/hmmufotu-sum.cpp:366-379; hmmufotu-jplace, src/hmmufotu-jplace.cpp:
// 226-241): the reference's own readProgInfo() on the text header, its TSVScanner with a header line, and every field those
// programs fetch BY NAME.  Prints one line per record, so that a TSV written by this repository's hmmufotu can be shown to
// be one they read, with the same values as the reference's own TSV.
#include <fstream>
#include <functional>
#include <iostream>
#include <string>
#include "HmmUFOtu_common.h"

using namespace std;
using namespace EGriceLab;
using namespace EGriceLab::HmmUFOtu;

int main(int argc, char** argv) {
	if (argc != 2) { cerr << "usage: tsv_probe ASSIGNMENT.tsv" << endl; return 2; }
	ifstream in(argv[1]);
	if (!in.is_open()) return 1;
	if (readProgInfo(in).bad()) { cerr << "readProgInfo rejected the header" << endl; return 1; }
	TSVScanner tsvIn(in, true);
	long n = 0;
	while (tsvIn.hasNext()) {
		const TSVRecord& r = tsvIn.nextRecord();
		cout << r.getFieldByName("id") << "\t" << ::atoi(r.getFieldByName("CS_start").c_str()) << "\t" << ::atoi(r.getFieldByName("CS_end").c_str())
		     << "\t" << hash<string>()(r.getFieldByName("alignment")) << "\t" << r.getFieldByName("branch_id") << "\t" << ::atof(r.getFieldByName("branch_ratio").c_str())
		     << "\t" << ::atol(r.getFieldByName("taxon_id").c_str()) << "\t" << ::atof(r.getFieldByName("anno_dist").c_str()) << "\t" << ::atof(r.getFieldByName("loglik").c_str())
		     << "\t" << ::atof(r.getFieldByName("Q_placement").c_str()) << "\t" << ::atof(r.getFieldByName("Q_taxon").c_str()) << "\n";
		n++;
	}
	cerr << n << " records" << endl;
	return 0;
}
