/* prints sizeof / offsetof of every structure of the dabsdr C API; compiled once against the reference's header
 * (tests/golden/make_abi_layout.py -> abi_layout.json) and, in tests/test_abi.py, against include/dabsdr_abi.h */
#include <stddef.h>
#include <stdio.h>
#include ABI_HEADER
#define S(t) printf("\"sizeof(" #t ")\": %zu,\n", sizeof(t))
#define O(t, f) printf("\"offsetof(" #t "," #f ")\": %zu,\n", offsetof(t, f))
int main(void)
{
    printf("{\n");
    S(dabsdrVersion_t); S(dabsdrAudioFrameHeader_t); S(dabsdrAudioCBData_t); O(dabsdrAudioCBData_t, ASCTy); O(dabsdrAudioCBData_t, header);
    O(dabsdrAudioCBData_t, auLen); O(dabsdrAudioCBData_t, pAuData);
    S(dabsdrDynamicLabelCBData_t); O(dabsdrDynamicLabelCBData_t, len); O(dabsdrDynamicLabelCBData_t, pData);
    S(dabsdrDataGroupCBData_t); O(dabsdrDataGroupCBData_t, SCId); O(dabsdrDataGroupCBData_t, userAppType); O(dabsdrDataGroupCBData_t, dgLen); O(dabsdrDataGroupCBData_t, pDgData);
    S(dabsdrNotificationCBData_t); O(dabsdrNotificationCBData_t, status); O(dabsdrNotificationCBData_t, len); O(dabsdrNotificationCBData_t, pData);
    S(dabsdrNtfSyncStatus_t); O(dabsdrNtfSyncStatus_t, snr10);
    S(dabsdrNtfPeriodic_t); O(dabsdrNtfPeriodic_t, snr10); O(dabsdrNtfPeriodic_t, freqOffset); O(dabsdrNtfPeriodic_t, dateHoursMinutes); O(dabsdrNtfPeriodic_t, secMsec);
    O(dabsdrNtfPeriodic_t, fibErrorCntr); O(dabsdrNtfPeriodic_t, mscCrcOkCntr); O(dabsdrNtfPeriodic_t, mscCrcErrorCntr); O(dabsdrNtfPeriodic_t, audioServiceBytes);
    O(dabsdrNtfPeriodic_t, padBytes); O(dabsdrNtfPeriodic_t, rsUncorrectableCntr); O(dabsdrNtfPeriodic_t, rsBitErrors); O(dabsdrNtfPeriodic_t, rsBytes);
    S(dabsdrLabel_t); O(dabsdrLabel_t, charField); O(dabsdrLabel_t, charset);
    S(dabsdrServiceListItem_t); O(dabsdrServiceListItem_t, label); O(dabsdrServiceListItem_t, pty); O(dabsdrServiceListItem_t, CAId);
    S(dabsdrServiceCompListItem_t); O(dabsdrServiceCompListItem_t, SubChId); O(dabsdrServiceCompListItem_t, SubChAddr); O(dabsdrServiceCompListItem_t, SubChSize);
    O(dabsdrServiceCompListItem_t, protectionLevel); O(dabsdrServiceCompListItem_t, uepIdx); O(dabsdrServiceCompListItem_t, ps); O(dabsdrServiceCompListItem_t, lang);
    O(dabsdrServiceCompListItem_t, CAflag); O(dabsdrServiceCompListItem_t, label); O(dabsdrServiceCompListItem_t, numUserApps); O(dabsdrServiceCompListItem_t, TMId);
    O(dabsdrServiceCompListItem_t, streamAudio.ASCTy); O(dabsdrServiceCompListItem_t, streamAudio.bitRate); O(dabsdrServiceCompListItem_t, packetData.SCId);
    O(dabsdrServiceCompListItem_t, packetData.DGflag); O(dabsdrServiceCompListItem_t, packetData.packetAddress);
    S(dabsdrUserAppListItem_t); O(dabsdrUserAppListItem_t, label); O(dabsdrUserAppListItem_t, dataLen); O(dabsdrUserAppListItem_t, data);
    S(dabsdrNtfServiceList_t); O(dabsdrNtfServiceList_t, getServiceListItem);
    S(dabsdrNtfServiceComponentList_t); O(dabsdrNtfServiceComponentList_t, numServiceComponents); O(dabsdrNtfServiceComponentList_t, getServiceComponentListItem);
    S(dabsdrNtfAnnouncementSupport_t); O(dabsdrNtfAnnouncementSupport_t, ASu); O(dabsdrNtfAnnouncementSupport_t, numClusterIds); O(dabsdrNtfAnnouncementSupport_t, clusterIds);
    S(dabsdrNtfUserAppUpdate_t); S(dabsdrNtfUserAppList_t); O(dabsdrNtfUserAppList_t, numUserApps); O(dabsdrNtfUserAppList_t, getUserAppListItem);
    S(dabsdrNtfEnsemble_t); O(dabsdrNtfEnsemble_t, ueid); O(dabsdrNtfEnsemble_t, LTO); O(dabsdrNtfEnsemble_t, intTable); O(dabsdrNtfEnsemble_t, alarm); O(dabsdrNtfEnsemble_t, label);
    S(dabsdrNtfServiceSelection_t); O(dabsdrNtfServiceSelection_t, SCIdS); O(dabsdrNtfServiceSelection_t, id);
    S(dabsdrNtfXpadAppStartStop_t); S(dabsdrAsw_t); S(dabsdrNtfAnnouncementSwitching_t); S(dabsdrNtfPTy_t);
    S(dabsdrTii_t); O(dabsdrTii_t, sub); O(dabsdrTii_t, level); S(dabsdrNtfTii_t); O(dabsdrNtfTii_t, id); O(dabsdrNtfTii_t, getSpectrumTii);
    printf("\"DABSDR_NID_TII\": %d, \"DABSDR_NID_PERIODIC\": %d, \"DABSDR_NID_RESET\": %d, \"DABSDR_NSTAT_SERVICE_NOT_SUPPORTED\": %d,\n",
           (int)DABSDR_NID_TII, (int)DABSDR_NID_PERIODIC, (int)DABSDR_NID_RESET, (int)DABSDR_NSTAT_SERVICE_NOT_SUPPORTED);
    printf("\"DABSDR_SYNC_LEVEL_FIC\": %d, \"DABSDR_ID_DATA\": %d, \"DABSDR_SPECT_TII\": %d, \"DABSDR_TII_MODE_DEFAULT\": %d\n}\n",
           (int)DABSDR_SYNC_LEVEL_FIC, (int)DABSDR_ID_DATA, (int)DABSDR_SPECT_TII, (int)DABSDR_TII_MODE_DEFAULT);
    return 0;
}
