// TEST INFRASTRUCTURE (built by oracle/Makefile.ref into oracle/_ref/csfm_probe): the reference's own CSFMIndex
// (src/CSFMIndex.{h,cpp} + libcds, compiled in place) loaded from a <db>.csfm file written by the reference's
// hmmufotu-build; prints, for every pattern on stdin, count() and the sorted locate() hits as
// "pattern<TAB>count<TAB>start,end,CS ..." so that the host index of this repository (rebuilt from <db>.msa)
// can be compared with it hit for hit.
#include <algorithm>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include "HmmUFOtu_common.h"
#include "HmmUFOtu_hmm.h"

using namespace std;
using namespace EGriceLab;
using namespace EGriceLab::HmmUFOtu;

int main(int argc, char** argv) {
	if (argc != 2) { cerr << "usage: csfm_probe DB.csfm < patterns" << endl; return 2; }
	ifstream in(argv[1], ios_base::binary);
	if (loadProgInfo(in).bad()) return 1;
	CSFMIndex csfm;
	csfm.load(in);
	if (in.bad()) { cerr << "cannot load " << argv[1] << endl; return 1; }
	string pat;
	while (getline(cin, pat)) {
		if (pat.empty()) continue;
		vector<CSLoc> locs = csfm.locate(pat);
		vector<string> items;
		for (size_t i = 0; i < locs.size(); ++i) {
			string s = to_string(locs[i].start) + "," + to_string(locs[i].end) + "," + locs[i].CS;
			items.push_back(s);
		}
		sort(items.begin(), items.end());
		cout << pat << "\t" << csfm.count(pat);
		for (size_t i = 0; i < items.size(); ++i) cout << "\t" << items[i];
		cout << "\n";
	}
	return 0;
}
