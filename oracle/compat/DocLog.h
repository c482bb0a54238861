// TEST INFRASTRUCTURE: in-memory stand-in for the reference's CDocLog (source/DocLog.h:27-67).
// Collects the lines the scan decoder emits so the harness can count error lines.
#pragma once
#include "mfc_stub.h"
class CDocLog {
public:
	std::vector<std::string> lines;   // every line
	std::vector<std::string> errs;    // AddLineErr lines only
	std::vector<std::string> warns;   // AddLineWarn lines only
	bool enabled = true;
	void AddLine(CString s)        { if(enabled) lines.push_back(s.s); }
	void AddLineHdr(CString s)     { if(enabled) lines.push_back(s.s); }
	void AddLineHdrDesc(CString s) { if(enabled) lines.push_back(s.s); }
	void AddLineWarn(CString s)    { if(enabled){ lines.push_back(s.s); warns.push_back(s.s);} }
	void AddLineErr(CString s)     { if(enabled){ lines.push_back(s.s); errs.push_back(s.s);} }
	void AddLineGood(CString s)    { if(enabled) lines.push_back(s.s); }
	void Enable()  { enabled = true; }
	void Disable() { enabled = false; }
	bool quick=false;
	void SetQuickMode(bool b){quick=b;} bool GetQuickMode(){return quick;}
	void Clear()   { lines.clear(); errs.clear(); warns.clear(); }
};
